"""Row (f-3): data/embeddings/<model>/ CSV layout as the reference's train_*_full.py scripts write it."""
import ast
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def test_embedding_and_prediction_files(tmp_path):
    from prob_matrix_factorization_b200 import io, synth
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    u, i, x = synth.make_ratings(200, 100, 2000, seed=3)
    df = synth.to_frame(u.astype(np.int64), i.astype(np.int64), x.astype(float))
    m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=6, max_iter=2, tol=None, verbose=False)).fit(df)
    out = io.save_embeddings(m, str(tmp_path), recipe_ids=np.arange(100) + 1000)
    assert out.endswith(os.path.join("embeddings", "poisson_mf"))
    ue = pd.read_csv(os.path.join(out, "user_embeddings.csv"), float_precision="round_trip")
    ie = pd.read_csv(os.path.join(out, "item_embeddings.csv"), float_precision="round_trip")
    assert list(ue.columns) == [str(k) for k in range(6)] and ue.shape == (200, 6)
    assert list(ie.columns) == ["recipe_id"] + [str(k) for k in range(6)] and ie.shape == (100, 7)
    assert np.array_equal(ue.to_numpy(), m.E_theta)           # shortest round-trip float repr, float64
    cfg = ast.literal_eval(open(os.path.join(out, "config.txt")).read())
    assert PoissonMFCAVIConfig(**cfg) == m.config
    pred = m.predict(u[:50], i[:50])
    pdir = io.save_test_predictions(m, u[:50], i[:50], x[:50], pred, str(tmp_path))
    tp = pd.read_csv(os.path.join(pdir, "test_predictions.csv"), float_precision="round_trip")
    assert list(tp.columns) == ["u", "i", "y_true", "y_pred"] and np.array_equal(tp.y_pred.to_numpy(), pred)
