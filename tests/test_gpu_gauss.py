"""a5 on the GPU: Gaussian MF CAVI (bias / no-bias) vs golden outputs of the reference and vs the oracle."""
import numpy as np
import pandas as pd
import pytest

from conftest import rel_max
from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def frame(u, i, x):
    return pd.DataFrame({"u": np.asarray(u, np.int64), "i": np.asarray(i, np.int64), "rating": np.asarray(x, float)})


def test_gaussian_bias_golden(golden):
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    g = golden("gaussian_bias")
    cfg = GaussianMFCAVIConfig(n_factors=g["K"], sigma2=g["sigma2"], eta_theta2=g["eta_theta2"], eta_beta2=g["eta_beta2"],
                               eta_bias2=g["eta_bias2"], max_iter=g["T"], tol=1e-3, random_state=g["seed"], verbose=False)
    m = GaussianMFCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]), global_mean=g["global_mean"])
    for k in ("m_theta", "V_theta", "m_beta", "V_beta", "m_user_bias", "m_item_bias"):
        got = getattr(m, k)
        assert got.shape == g[k].shape and got.dtype == np.float64
        assert rel_max(got, g[k]) < TOL, k
    val = frame(g["val_u"], g["val_i"], g["val_x"])
    assert rel_max(m.predict(g["val_u"], g["val_i"], g["global_mean"]), g["val_pred"]) < TOL
    assert abs(m.evaluate_rmse(val, g["global_mean"]) - g["val_rmse"]) < TOL * g["val_rmse"]
    assert abs(m.evaluate_macro_mae(val, g["global_mean"]) - g["val_macro_mae"]) < TOL * g["val_macro_mae"]
    test = frame(g["test_u"], g["test_i"], g["test_x"])
    assert abs(m.evaluate_rmse(test, g["global_mean"]) - g["test_rmse"]) < TOL * g["test_rmse"]
    # GaussianLogPredictiveLikelihood (metrics.py:18-35; the golden value was computed by the reference's function on the
    # rows with seen ids, with sigma2 passed as `sigma` as the reference's scripts do) -- device path and host function
    from prob_matrix_factorization_b200 import metrics
    assert abs(m.log_predictive_likelihood(test) - g["test_lpl"]) < TOL * abs(g["test_lpl"])
    seen = test[(test.u < m.n_users) & (test.i < m.n_items)]
    assert abs(metrics.GaussianLogPredictiveLikelihood(seen, m.m_theta, m.m_beta, g["sigma2"]) - g["test_lpl"]) < TOL * abs(g["test_lpl"])
    # early stopping needs 0 <= improvement < tol (gaussian_mf_cavi_bias.py:279)
    cfg2 = GaussianMFCAVIConfig(n_factors=g["K"], sigma2=g["sigma2"], eta_theta2=g["eta_theta2"], eta_beta2=g["eta_beta2"],
                                eta_bias2=g["eta_bias2"], max_iter=g["es_max_iter"], tol=g["es_tol"],
                                random_state=g["seed"], verbose=False)
    m2 = GaussianMFCAVI(cfg2).fit(frame(g["u"], g["i"], g["x"]), val, global_mean=g["global_mean"])
    assert m2.n_iter_ == g["es_iterations"]
    assert rel_max(m2.m_theta, g["es_m_theta"]) < TOL


def test_gaussian_nobias_golden(golden):
    from prob_matrix_factorization_b200.gaussian_mf_cavi import GaussianMFCAVI, GaussianMFCAVIConfig
    g = golden("gaussian_nobias")
    cfg = GaussianMFCAVIConfig(n_factors=g["K"], sigma2=g["sigma2"], eta_theta2=g["eta_theta2"], eta_beta2=g["eta_beta2"],
                               max_iter=g["T"], tol=1e-3, random_state=g["seed"], verbose=False)
    m = GaussianMFCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]), global_mean=g["global_mean"])
    for k in ("m_theta", "V_theta", "m_beta", "V_beta"):
        assert rel_max(getattr(m, k), g[k]) < TOL, k
    assert rel_max(m.predict(g["val_u"], g["val_i"], g["global_mean"]), g["val_pred"]) < TOL
    assert abs(m.evaluate_rmse(frame(g["val_u"], g["val_i"], g["val_x"]), g["global_mean"]) - g["val_rmse"]) < TOL * g["val_rmse"]
    assert not hasattr(m, "m_user_bias")


@pytest.mark.parametrize("K,seg_len,bias", [(10, 64, True), (30, 8, True), (17, 16, False), (64, 64, True)])
def test_gaussian_vs_oracle(K, seg_len, bias):
    """Larger K (multi-slot threads), rows cut into many segments, heavy rows; 5 sweeps."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    from prob_matrix_factorization_b200 import gaussian_mf_cavi as NB
    N, M, nnz, T = 500, 300, 12_000, 5
    u, i, x = synth.make_ratings(N, M, nnz, seed=300 + K)
    x = x.astype(np.float64) - x.mean()
    hp = dict(sigma2=0.3, eta_theta2=0.5, eta_beta2=0.5)
    if bias:
        m = GaussianMFCAVI(GaussianMFCAVIConfig(n_factors=K, eta_bias2=1.0, max_iter=T, verbose=False, **hp), seg_len=seg_len)
    else:
        m = NB.GaussianMFCAVI(NB.GaussianMFCAVIConfig(n_factors=K, max_iter=T, verbose=False, **hp), seg_len=seg_len)
    m.fit(frame(u, i, x))
    ref = O.gauss_sweeps(u, i, x, K, 0.3, 0.5, 0.5, 1.0, T, 42, bias=bias, n_users=m.n_users, n_items=m.n_items)
    keys = ["m_theta", "V_theta", "m_beta", "V_beta"] + (["m_user_bias", "m_item_bias"] if bias else [])
    for k in keys:
        assert rel_max(getattr(m, k), ref[k]) < TOL, (k, K)


@pytest.mark.parametrize("bias", [True, False])
def test_gaussian_c1_full_size(bias):
    """BASELINE configs[0] at its real shape (gaussian_mf K=10, 20k users x 10k recipes x 200k ratings, hyper-parameters of
    best_hyperparams.txt:3), 20 sweeps, against the oracle's C port (pinned on the reference's outputs at 1e-10)."""
    from oracle import c_oracle as CO
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    from prob_matrix_factorization_b200 import gaussian_mf_cavi as NB
    w, (u, i, x) = synth.workload_ratings("c1")
    K, T = w.n_factors, 20
    mean = float(x.mean())
    xc = x.astype(np.float64) - mean                       # compare_models.py:54-58: the caller centres the ratings
    hp = dict(sigma2=0.5, eta_theta2=0.1, eta_beta2=0.1)
    if bias:
        m = GaussianMFCAVI(GaussianMFCAVIConfig(n_factors=K, eta_bias2=0.1, max_iter=T, random_state=42, verbose=False, **hp))
    else:
        m = NB.GaussianMFCAVI(NB.GaussianMFCAVIConfig(n_factors=K, max_iter=T, random_state=42, verbose=False, **hp))
    m.fit(frame(u, i, xc), global_mean=mean)
    assert (m.n_users, m.n_items) == (w.n_users, w.n_items)
    init = O.gauss_init(w.n_users, w.n_items, K, 42)
    ref = CO.gauss_sweeps(u, i, xc, w.n_users, w.n_items, K, 0.5, 0.1, 0.1, 0.1, T, init, bias=bias)
    keys = ["m_theta", "V_theta", "m_beta", "V_beta"] + (["m_user_bias", "m_item_bias"] if bias else [])
    for k in keys:
        assert rel_max(getattr(m, k), ref[k]) < TOL, k
