"""Boundary (b), script level: a reference-style driver (tests/ref_style/train_poisson_like.py -- the flow of
train_poisson_full.py:15-123: CSV splits -> hyper-parameters from a dict-literal file -> Config(**dict) -> fit ->
embeddings / config / predictions CSVs) runs UNCHANGED on top of ``dropin`` (its ``src.models.*`` /
``src.evaluation.metrics`` imports resolve to this engine, ``matplotlib`` is stubbed) in a fresh process, and what it
writes equals the oracle's result for the same inputs."""
import ast
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from conftest import REPO, rel_max
from oracle import c_oracle as CO
from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu


def test_reference_style_driver_runs_on_the_dropin(tmp_path):
    from prob_matrix_factorization_b200 import synth
    N, M, nnz, K, T = 1500, 900, 30_000, 12, 15
    tr, va, te = synth.make_splits(N, M, nnz, seed=404)
    os.makedirs(tmp_path / "data" / "processed")
    for name, (u, i, x) in (("train", tr), ("validation", va), ("test", te)):
        synth.to_frame(u, i, x).to_csv(tmp_path / "data" / "processed" / f"interactions_{name}.csv", index=False)
    hp = {"n_factors": K, "a0": 0.1, "b0": 0.5, "max_iter": T, "tol": None, "random_state": 42, "verbose": False}
    (tmp_path / "best_hyperparams.txt").write_text(f"GaussianMF: {{'n_factors': 3}}\nPoissonMF: {hp}\n")
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "prob_matrix_factorization_b200.dropin", "--stub-matplotlib", "--script",
                        os.path.join(REPO, "tests", "ref_style", "train_poisson_like.py")],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test Set Metrics: MacroMAE=" in r.stdout
    # what the driver wrote, against the oracle on the same (train + validation) ratings
    u = np.concatenate([tr[0], va[0]]); i = np.concatenate([tr[1], va[1]]); x = np.concatenate([tr[2], va[2]])
    n_users, n_items = int(u.max()) + 1, int(i.max()) + 1
    init = O.poisson_init(n_users, n_items, K, 0.1, 0.5, 42)
    ref = CO.poisson_sweeps(u, i, x, n_users, n_items, K, 0.1, 0.5, T, init["E_theta"], init["E_beta"])
    emb = tmp_path / "data" / "embeddings" / "poisson_mf"
    ue, ie = pd.read_csv(emb / "user_embeddings.csv"), pd.read_csv(emb / "item_embeddings.csv")
    assert list(ue.columns) == [str(k) for k in range(K)] and ue.shape == (n_users, K) and ie.shape == (n_items, K)
    assert rel_max(ue.to_numpy(), ref["E_theta"]) < 1e-5 and rel_max(ie.to_numpy(), ref["E_beta"]) < 1e-5
    assert ast.literal_eval((emb / "config.txt").read_text()) == hp
    pred = pd.read_csv(tmp_path / "data" / "predictions" / "poisson_mf" / "test_predictions.csv")
    assert list(pred.columns) == ["u", "i", "y_true", "y_pred"] and len(pred) == len(te[0])
    want = O.predict(te[0].astype(np.int64), te[1].astype(np.int64), ref["E_theta"], ref["E_beta"])
    assert rel_max(pred["y_pred"].to_numpy(), want) < 1e-5
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("Test Set Metrics")][0]
    assert f"RMSE={O.rmse(te[2].astype(float), want):.4f}" in line
