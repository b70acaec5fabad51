"""a3/a4 + a8-a10 on the GPU: Poisson MF and HPF-CAVI through the drop-in classes (C-ABI underneath)
against (i) golden outputs of the reference itself, (ii) the oracle at sizes it finishes in seconds.

Tolerance (BASELINE.json north_star): float32 engine vs float64 reference, max-norm relative
error ||d||inf/||ref||inf <= 1e-5 after a fixed number of sweeps; integer outputs bit-exact.
"""
import numpy as np
import pandas as pd
import pytest

from conftest import rel_max
from oracle import c_oracle as CO
from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def frame(u, i, x):
    return pd.DataFrame({"u": np.asarray(u, np.int64), "i": np.asarray(i, np.int64), "rating": np.asarray(x, float)})


def test_poisson_golden(golden):
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    g = golden("poisson")
    cfg = PoissonMFCAVIConfig(n_factors=g["K"], a0=g["a0"], b0=g["b0"], max_iter=g["T"], tol=None,
                              random_state=g["seed"], verbose=False)
    m = PoissonMFCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]))
    assert (m.n_users, m.n_items) == (g["n_users"], g["n_items"])
    for k in ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta"):
        got = getattr(m, k)
        assert got.dtype == np.float64 and got.shape == g[k].shape
        assert rel_max(got, g[k]) < TOL, k
    val = frame(g["val_u"], g["val_i"], g["val_x"])
    pred = m.predict(g["val_u"], g["val_i"])
    assert pred.dtype == np.float64 and rel_max(pred, g["val_pred"]) < TOL
    assert (pred[:5] == 0).all()                       # unseen ids -> 0, still scored
    assert abs(m.evaluate_rmse(val) - g["val_rmse"]) < TOL * g["val_rmse"]
    assert abs(m.evaluate_macro_mae(val) - g["val_macro_mae"]) < TOL * g["val_macro_mae"]
    test = frame(g["test_u"], g["test_i"], g["test_x"])
    assert abs(m.evaluate_rmse(test) - g["test_rmse"]) < TOL * g["test_rmse"]
    assert abs(m.log_predictive_likelihood(test) - g["test_lpl"]) < 1e-5 * abs(g["test_lpl"])


def test_poisson_early_stopping_matches_reference(golden):
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    g = golden("poisson")
    cfg = PoissonMFCAVIConfig(n_factors=g["K"], a0=g["a0"], b0=g["b0"], max_iter=g["es_max_iter"], tol=g["es_tol"],
                              random_state=g["seed"], verbose=False)
    m = PoissonMFCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]), frame(g["val_u"], g["val_i"], g["val_x"]))
    assert m.n_iter_ == g["es_iterations"]
    assert rel_max(m.E_theta, g["es_E_theta"]) < TOL
    assert abs(m.val_rmse_history_[-1] - g["es_val_rmse"]) < TOL * g["es_val_rmse"]


def test_hpf_golden(golden):
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    g = golden("hpf_cavi")
    hp = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    cfg = HPF_CAVI_Config(n_factors=g["K"], max_iter=g["T"], tol=None, random_state=g["seed"], verbose=False, **hp)
    m = HPF_CAVI(cfg).fit(frame(g["u"], g["i"], g["x"]))
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
              "E_theta", "E_beta", "E_xi", "E_eta"):
        assert rel_max(getattr(m, k), g[k]) < TOL, k
    assert m.gamma_a_xi == g["gamma_a_xi"] and m.gamma_a_eta == g["gamma_a_eta"]
    val = frame(g["val_u"], g["val_i"], g["val_x"])
    assert rel_max(m.predict(g["val_u"], g["val_i"]), g["val_pred"]) < TOL
    assert abs(m.evaluate_rmse(val) - g["val_rmse"]) < TOL * g["val_rmse"]
    assert abs(m.evaluate_macro_mae(val) - g["val_macro_mae"]) < TOL * g["val_macro_mae"]
    cfg2 = HPF_CAVI_Config(n_factors=g["K"], max_iter=g["es_max_iter"], tol=g["es_tol"], random_state=g["seed"],
                           verbose=False, **hp)
    m2 = HPF_CAVI(cfg2).fit(frame(g["u"], g["i"], g["x"]), val)
    assert m2.n_iter_ == g["es_iterations"]
    assert rel_max(m2.E_theta, g["es_E_theta"]) < TOL


@pytest.mark.parametrize("K,seg_len", [(3, 8), (8, 64), (10, 128), (20, 16), (50, 128), (64, 64), (70, 32), (100, 128), (130, 64)])
def test_poisson_vs_oracle_sizes(K, seg_len):
    """Every (G,V) kernel variant, rows cut into many segments, 20 sweeps (T=20)."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    N, M, nnz, T = 3000, 1500, 60_000, 20
    u, i, x = synth.make_ratings(N, M, nnz, seed=100 + K)
    cfg = PoissonMFCAVIConfig(n_factors=K, a0=0.1, b0=0.5, max_iter=T, tol=None, random_state=7, verbose=False)
    m = PoissonMFCAVI(cfg, seg_len=seg_len)
    m.n_users, m.n_items = N, M
    init = m._initial_state()
    m.fit_arrays(u, i, x, init)
    ref = CO.poisson_sweeps(u, i, x, N, M, K, 0.1, 0.5, T, init["E_theta"], init["E_beta"])
    for k in ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta"):
        assert rel_max(getattr(m, k), ref[k]) < TOL, (k, K)


def test_hpf_vs_oracle_c1_shape():
    """BASELINE config-1 shape (20k x 10k x 200k) with the HPF model, K=50, 20 sweeps."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    N, M, nnz, K, T = 20_000, 10_000, 200_000, 50, 20
    u, i, x = synth.make_ratings(N, M, nnz, seed=20261)
    x = x + 1.0
    hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
    m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=None, random_state=42, verbose=False, **hp))
    m.n_users, m.n_items = N, M
    init = m._initial_state()
    m.fit_arrays(u, i, x, init)
    ref = CO.hpf_sweeps(u, i, x, N, M, K, hp, T, init)
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
              "E_theta", "E_beta", "E_xi", "E_eta"):
        assert rel_max(getattr(m, k), ref[k]) < TOL, k


def test_zero_iterations_and_rate_floor():
    """max_iter=0 exposes the initial state; all-zero factors hit the 1e-10 clamp (poisson_mf_cavi.py:153)."""
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    u = np.array([0, 0, 1, 2, 2, 2]); i = np.array([0, 1, 1, 0, 2, 2]); x = np.array([3, 0, 5, 1, 2, 2.0])
    m0 = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=4, max_iter=0, verbose=False)).fit(frame(u, i, x))
    st = O.poisson_init(3, 3, 4, 0.3, 1.0, 42)
    assert rel_max(m0.E_theta, st["E_theta"]) < 1e-6 and rel_max(m0.a_theta, st["a_theta"]) < 1e-12
    m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=4, a0=0.3, b0=1.0, max_iter=1, verbose=False))
    m.n_users, m.n_items = 3, 3
    init = m._initial_state()
    init["E_theta"] = np.zeros_like(init["E_theta"]); init["E_beta"] = np.zeros_like(init["E_beta"])
    m.fit_arrays(u, i, x, init)
    ref = CO.poisson_sweeps(u, i, x, 3, 3, 4, 0.3, 1.0, 1, init["E_theta"], init["E_beta"])
    assert rel_max(m.E_theta, ref["E_theta"]) < TOL and rel_max(m.E_beta, ref["E_beta"]) < TOL


@pytest.mark.parametrize("model", ["poisson", "hpf"])
@pytest.mark.parametrize("user_tiles,item_tiles,seg_len", [(1, 4, 64), (3, 1, 16), (3, 5, 8), (2, 2, 128)])
def test_tiled_passes_vs_oracle(model, user_tiles, item_tiles, seg_len):
    """The tiled form of the passes (ratings split by the id range of the gathered table, row sums carried across the
    tiles by pmf_gamma_pass_acc): same results as the untiled reference algorithm.  seg_len 8/16 puts multi-segment
    rows (the scratch + gamma_multi_kernel path) into every tile; rows without ratings in a tile exercise the
    skip / zero / finish-from-running-sums branches."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
    N, M, nnz, K, T = 4000, 2500, 50_000, 20, 10
    u, i, x = synth.make_ratings(N, M, nnz, seed=321)
    hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
    if model == "poisson":
        m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=K, a0=0.1, b0=0.5, max_iter=T, tol=None, verbose=False), seg_len=seg_len)
    else:
        x = x + 1.0
        m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=None, verbose=False, **hp), seg_len=seg_len)
    m._ratings_kw = dict(user_pass_tiles=user_tiles, item_pass_tiles=item_tiles)
    m.n_users, m.n_items = N, M
    init = m._initial_state()
    m.fit_arrays(u, i, x, init)
    assert len(m._engine.r.user_tiles) == user_tiles and len(m._engine.r.item_tiles) == item_tiles
    assert sum(g.nnz for g in m._engine.r.user_tiles) == nnz and sum(g.nnz for g in m._engine.r.item_tiles) == nnz
    if model == "poisson":
        ref = CO.poisson_sweeps(u, i, x, N, M, K, 0.1, 0.5, T, init["E_theta"], init["E_beta"])
        names = ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta")
    else:
        ref = CO.hpf_sweeps(u, i, x, N, M, K, hp, T, init)
        names = ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
                 "E_theta", "E_beta", "E_xi", "E_eta")
    for k in names:
        assert rel_max(getattr(m, k), ref[k]) < TOL, k


def test_coo_partition_is_stable():
    """pmf_coo_partition (routing of ratings to tiles / owner ranks): bit-exact stable bucket partition."""
    import torch
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.ratings import coo_partition
    N, M, nnz = 5000, 3000, 200_001
    u, i, x = synth.make_ratings(N, M, nnz, seed=5)
    ud, idv, xd = (torch.from_numpy(a).cuda() for a in (u, i, x))
    for by_item, ids, bounds in ((False, u, [0, 17, 17, 2500, 4999, 5000]), (True, i, [0, 3000]), (True, i, [0, 1, 2, 1000, 3000])):
        uo, io, xo, offs = coo_partition(ud, idv, xd, np.array(bounds), by_item)
        bucket = np.searchsorted(np.array(bounds[1:]), ids, side="right")
        order = np.argsort(bucket, kind="stable")
        assert np.array_equal(uo.cpu().numpy(), u[order]) and np.array_equal(io.cpu().numpy(), i[order])
        assert np.array_equal(xo.cpu().numpy(), x[order])
        assert np.array_equal(offs, np.concatenate([[0], np.cumsum(np.bincount(bucket, minlength=len(bounds) - 1))]))


def test_count_keys():
    import torch
    from prob_matrix_factorization_b200 import _cabi, synth
    u, _, _ = synth.make_ratings(7000, 10, 100_003, seed=9)
    ud = torch.from_numpy(u).cuda()
    counts = torch.full((7000,), -1, dtype=torch.int32, device="cuda")
    _cabi.call("pmf_count_keys", ud.data_ptr(), ud.numel(), 7000, counts.data_ptr(), _cabi.stream_ptr())
    assert np.array_equal(counts.cpu().numpy(), np.bincount(u, minlength=7000))
    with pytest.raises(_cabi.PMFError):
        _cabi.call("pmf_count_keys", ud.data_ptr(), ud.numel(), 6999, counts.data_ptr(), _cabi.stream_ptr())


@pytest.mark.parametrize("tol", [None, 2e-3])
def test_device_loop_equals_host_loop(tol, monkeypatch):
    """f2: the fit loop as one CUDA graph with a device-side WHILE (sweeps + validation statistics + the reference's
    stopping rule, hpf_cavi.py:196-211) stops at the same iteration, with the same RMSE history and bit-identical
    factors as the host loop that reads the statistics back after every sweep."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
    N, M, nnz, K = 3000, 2000, 40_000, 12
    (u, i, x), (vu, vi, vx), _ = synth.make_splits(N, M, nnz, seed=11)
    hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PMF_DEVICE_LOOP", mode)
        m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=25, tol=tol, random_state=3, verbose=False, **hp))
        m.fit(frame(u, i, x + 1.0), frame(vu, vi, vx + 1.0))
        out[mode] = (m.n_iter_, np.array(m.val_rmse_history_), m.E_theta, m.gamma_b_beta, m.E_eta)
    assert out["1"][0] == out["0"][0] and (tol is None or 2 <= out["1"][0] < 25)
    assert np.array_equal(out["1"][1], out["0"][1]) and len(out["1"][1]) == out["1"][0]
    for a, b in zip(out["1"][2:], out["0"][2:]):
        assert np.array_equal(a, b)
