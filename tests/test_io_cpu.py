"""Row (f-3) on CPU: the CSV layout of data/embeddings/<model>/ and data/predictions/<model>/ as pandas writes it in the
reference's train_*_full.py scripts (header 0..K-1, shortest round-trip floats, optional leading recipe_id, config.txt =
str(asdict(config)) [+ global_mean line]).  Stand-in model objects: the writers only read public arrays."""
import ast
import os
from dataclasses import dataclass

import numpy as np
import pandas as pd


@dataclass
class _Cfg:
    n_factors: int = 3
    tol: float = None


def _model(name, **arrays):
    cls = type(name, (), {})
    m = cls()
    m.config = _Cfg()
    for k, v in arrays.items():
        setattr(m, k, v)
    return m


def test_embeddings_layout_poisson_and_gaussian(tmp_path):
    from prob_matrix_factorization_b200 import io
    rng = np.random.default_rng(0)
    Et, Eb = rng.random((5, 3)), rng.random((4, 3))
    out = io.save_embeddings(_model("PoissonMFCAVI", E_theta=Et, E_beta=Eb), str(tmp_path), recipe_ids=[11, 12, 13, 14])
    assert out == os.path.join(str(tmp_path), "embeddings", "poisson_mf")
    ue = pd.read_csv(os.path.join(out, "user_embeddings.csv"), float_precision="round_trip")
    ie = pd.read_csv(os.path.join(out, "item_embeddings.csv"), float_precision="round_trip")
    assert list(ue.columns) == ["0", "1", "2"] and np.array_equal(ue.to_numpy(), Et)
    assert list(ie.columns) == ["recipe_id", "0", "1", "2"] and np.array_equal(ie.iloc[:, 1:].to_numpy(), Eb)
    assert ast.literal_eval(open(os.path.join(out, "config.txt")).read()) == {"n_factors": 3, "tol": None}
    # Gaussian: means are the embeddings, config.txt carries the global mean on a second line (train_gaussian_full.py:106)
    out = io.save_embeddings(_model("GaussianMFCAVI", m_theta=Et, m_beta=Eb), str(tmp_path), recipe_ids=[1, 2],   # wrong length
                             global_mean=4.25)
    ie = pd.read_csv(os.path.join(out, "item_embeddings.csv"))
    assert list(ie.columns) == ["0", "1", "2"]                      # ids of the wrong length are not attached
    lines = open(os.path.join(out, "config.txt")).read().split("\n")
    assert ast.literal_eval(lines[0]) == {"n_factors": 3, "tol": None} and lines[1] == "global_mean: 4.25"


def test_predictions_layout(tmp_path):
    from prob_matrix_factorization_b200 import io
    y = np.array([0.1, 2.5, 1 / 3])
    d = io.save_test_predictions(_model("HPF_CAVI"), [0, 1, 2], [5, 6, 7], [1.0, 2.0, 3.0], y, str(tmp_path))
    tp = pd.read_csv(os.path.join(d, "test_predictions.csv"), float_precision="round_trip")
    assert d.endswith(os.path.join("predictions", "hpf_cavi"))
    assert list(tp.columns) == ["u", "i", "y_true", "y_pred"] and np.array_equal(tp.y_pred.to_numpy(), y)
