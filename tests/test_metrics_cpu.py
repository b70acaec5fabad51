"""Host metric helpers (prob_matrix_factorization_b200/metrics.py, the drop-in for src/evaluation/metrics.py:6-66)
against values the reference's own functions produced (tests/golden/*.npz, written by make_golden.py) and against the
oracle's restatement."""
import numpy as np
import pandas as pd
import pytest

from oracle import pmf_oracle as O
from prob_matrix_factorization_b200 import metrics


def frame(u, i, x):
    return pd.DataFrame({"u": np.asarray(u, np.int64), "i": np.asarray(i, np.int64), "rating": np.asarray(x, float)})


def test_rmse_mae_macro_mae_match_the_reference(golden):
    g = golden("poisson")
    y, pred = g["val_x"].astype(float), g["val_pred"]
    assert metrics.rmse(y, pred) == pytest.approx(g["val_rmse"], rel=1e-12)
    assert metrics.macro_mae(y, pred) == pytest.approx(g["val_macro_mae"], rel=1e-12)
    assert metrics.mae(y, pred) == pytest.approx(np.mean(np.abs(y - pred)), rel=1e-15)
    assert metrics.rmse(y, pred) == pytest.approx(O.rmse(y, pred), rel=1e-15)
    assert metrics.macro_mae(y, pred) == pytest.approx(O.macro_mae(y, pred), rel=1e-15)


def test_macro_mae_weights_labels_equally():
    y = np.array([1.0, 1.0, 1.0, 5.0])
    p = np.array([1.0, 2.0, 3.0, 1.0])
    assert metrics.macro_mae(y, p) == pytest.approx(0.5 * (1.0 + 4.0))      # mean over labels {1, 5} of the per-label MAE
    assert metrics.mae(y, p) == pytest.approx(7.0 / 4.0)


def test_poisson_lpl_matches_the_reference(golden):
    g = golden("poisson")
    te = frame(g["test_u"], g["test_i"], g["test_x"])
    ok = (te.u < g["n_users"]) & (te.i < g["n_items"])
    # the golden value was computed by the reference's PoissonLogPredictiveLikelihood on the frame make_golden.py built
    got = metrics.PoissonLogPredictiveLikelihood(te[ok] if not ok.all() else te, g["E_theta"], g["E_beta"])
    assert got == pytest.approx(g["test_lpl"], rel=1e-10)
    assert got == pytest.approx(O.poisson_lpl(te[ok].u.to_numpy(), te[ok].i.to_numpy(), te[ok].rating.to_numpy(), g["E_theta"], g["E_beta"]), rel=1e-12)


def test_gaussian_lpl_squares_sigma_like_the_reference(golden):
    g = golden("gaussian_bias")
    te = frame(g["test_u"], g["test_i"], g["test_x"])
    seen = te[(te.u < g["n_users"]) & (te.i < g["n_items"])]
    got = metrics.GaussianLogPredictiveLikelihood(seen, g["m_theta"], g["m_beta"], g["sigma2"])
    assert got == pytest.approx(g["test_lpl"], rel=1e-10)
    # the quirk itself (metrics.py:33): the argument is squared although callers pass a variance
    pred = np.sum(g["m_theta"][seen.u] * g["m_beta"][seen.i], axis=1)
    v = g["sigma2"] ** 2
    assert got == pytest.approx(np.sum(-0.5 * np.log(2 * np.pi * v) - (seen.rating - pred) ** 2 / (2 * v)), rel=1e-12)
