"""§8f-4 on the GPU: the extended Poisson model (per-user phi, per-item psi) through its drop-in class against (i) golden
outputs of the reference itself and (ii) the oracle on a larger seeded problem with rows cut into several segments.

Tolerance: float32 engine vs float64 reference, max-norm relative error <= 1e-5 (BASELINE.json north_star).
"""
import numpy as np
import pandas as pd
import pytest

from conftest import rel_max
from oracle import pmf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
NAMES = ("a_theta", "b_theta", "a_beta", "b_beta", "a_phi", "b_phi", "a_psi", "b_psi", "E_theta", "E_beta", "E_phi", "E_psi")


def frame(u, i, x):
    return pd.DataFrame({"u": np.asarray(u, np.int64), "i": np.asarray(i, np.int64), "rating": np.asarray(x, float)})


def test_poisson_ext_golden(golden):
    from prob_matrix_factorization_b200.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig
    g = golden("poisson_ext")
    cfg = PoissonMFExtendedCAVIConfig(n_factors=g["K"], a0=g["a0"], b0=g["b0"], max_iter=g["T"], tol=None,
                                      random_state=g["seed"], verbose=False)
    m = PoissonMFExtendedCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]))
    assert (m.n_users, m.n_items) == (g["n_users"], g["n_items"])
    for k in NAMES:
        got = getattr(m, k)
        assert got.dtype == np.float64 and got.shape == g[k].shape, k
        assert rel_max(got, g[k]) < TOL, k
    pred = m.predict(g["val_u"], g["val_i"])
    assert pred.dtype == np.float64 and rel_max(pred, g["val_pred"]) < TOL
    assert (pred[:5] == 0).all()                                  # unseen ids -> 0
    assert abs(m.evaluate_rmse(frame(g["val_u"], g["val_i"], g["val_x"])) - g["val_rmse"]) < TOL * g["val_rmse"]
    assert abs(m.evaluate_rmse(frame(g["test_u"], g["test_i"], g["test_x"])) - g["test_rmse"]) < TOL * g["test_rmse"]


def test_poisson_ext_early_stopping_matches_reference(golden):
    from prob_matrix_factorization_b200.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig
    g = golden("poisson_ext")
    cfg = PoissonMFExtendedCAVIConfig(n_factors=g["K"], a0=g["a0"], b0=g["b0"], max_iter=g["es_max_iter"], tol=g["es_tol"],
                                      random_state=g["seed"], verbose=False)
    m = PoissonMFExtendedCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]), frame(g["val_u"], g["val_i"], g["val_x"]))
    assert m.n_iter_ == g["es_iterations"]
    assert rel_max(m.E_theta, g["es_E_theta"]) < TOL and rel_max(m.E_phi, g["es_E_phi"]) < TOL


def test_poisson_ext_before_any_sweep(golden):
    from prob_matrix_factorization_b200.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig
    g = golden("poisson_ext")
    cfg = PoissonMFExtendedCAVIConfig(n_factors=g["K"], a0=g["a0"], b0=g["b0"], max_iter=0, random_state=g["seed"], verbose=False)
    m = PoissonMFExtendedCAVI(cfg).fit(frame(g["u"], g["i"], g["x"]))
    init = O.poisson_ext_init(g["n_users"], g["n_items"], g["K"], g["a0"], g["b0"], g["seed"])
    for k in NAMES:
        assert rel_max(getattr(m, k), init[k]) < 1e-7, k


@pytest.mark.parametrize("K,seg_len", [(50, 64), (16, 64), (100, 128)])
def test_poisson_ext_vs_oracle_long_rows(K, seg_len):
    """Power-law rows: the heaviest rows span many segments (partial sums + rating sums combined by the second kernel),
    many rows are empty (prior shape/rate, untouched expectations)."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.poisson_mf_extended_cavi import PoissonMFExtendedCAVI, PoissonMFExtendedCAVIConfig
    N, M, nnz, T = 3000, 1500, 60_000, 4
    u, i, x = synth.make_ratings(N, M, nnz, 4321)
    x = x + 1.0                                                    # strictly positive: a zero dot product would be 0/0
    cfg = PoissonMFExtendedCAVIConfig(n_factors=K, a0=0.3, b0=1.0, max_iter=T, tol=None, random_state=7, verbose=False)
    m = PoissonMFExtendedCAVI(cfg, seg_len=seg_len).fit(frame(u, i, x))
    assert m._engine.r.by_user.n_multi_rows > 0 and m._engine.r.by_item.n_multi_rows > 0
    st = O.poisson_ext_sweeps(u, i, x, K, 0.3, 1.0, T, 7)
    for k in NAMES:
        assert rel_max(getattr(m, k), st[k]) < TOL, k
    rng = np.random.default_rng(1)
    pu, pi = rng.integers(0, N + 5, 4000), rng.integers(0, M + 5, 4000)
    assert rel_max(m.predict(pu, pi), O.poisson_ext_predict(pu, pi, st)) < TOL
