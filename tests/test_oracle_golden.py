"""Pin the CPU oracle: every restated function vs outputs of the reference's own classes.

The .npz files were produced by tests/golden/make_golden.py running the unmodified
reference (imported from /root/reference in the build container).  CPU only.
"""
import numpy as np
import pytest

from conftest import rel_max
from oracle import pmf_oracle as O

TIGHT = 1e-12  # float64 restatement vs float64 reference


def test_grouping_matches_reference_lists(golden):
    g = golden("poisson")
    for ids, n, perm_key, cnt_key in ((g["u"], g["n_users"], "user_perm", "user_counts"),
                                      (g["i"], g["n_items"], "item_perm", "item_counts")):
        row_ptr, perm = O.group_observations(ids, n)
        assert np.array_equal(perm, g[perm_key])
        assert np.array_equal(np.diff(row_ptr), g[cnt_key])
    assert (g["user_counts"] == 0).any() and (g["item_counts"] == 0).any(), "fixture must cover empty rows"


def test_poisson_sweeps(golden):
    g = golden("poisson")
    st = O.poisson_sweeps(g["u"], g["i"], g["x"], g["K"], g["a0"], g["b0"], g["T"], g["seed"])
    assert (st["n_users"], st["n_items"]) == (g["n_users"], g["n_items"])
    for k in ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta"):
        assert rel_max(st[k], g[k]) < TIGHT, k
    pred = O.predict(g["val_u"], g["val_i"], st["E_theta"], st["E_beta"])
    assert rel_max(pred, g["val_pred"]) < TIGHT
    assert (pred[:5] == 0).all()  # unseen ids predict 0 and still count (poisson_mf_cavi.py:228-241)
    assert abs(O.rmse(g["val_x"], pred) - g["val_rmse"]) < 1e-12
    assert abs(O.macro_mae(g["val_x"], pred) - g["val_macro_mae"]) < 1e-12
    lpl = O.poisson_lpl(g["test_u"], g["test_i"], g["test_x"], st["E_theta"], st["E_beta"])
    assert abs(lpl - g["test_lpl"]) < 1e-9 * abs(g["test_lpl"])


def test_poisson_ext_sweeps(golden):
    """Extended Poisson MF incl. its quirk: rows without observations keep their initial expectations."""
    g = golden("poisson_ext")
    st = O.poisson_ext_sweeps(g["u"], g["i"], g["x"], g["K"], g["a0"], g["b0"], g["T"], g["seed"])
    assert (st["n_users"], st["n_items"]) == (g["n_users"], g["n_items"])
    for k in ("a_theta", "b_theta", "a_beta", "b_beta", "a_phi", "b_phi", "a_psi", "b_psi", "E_theta", "E_beta", "E_phi", "E_psi"):
        assert rel_max(st[k], g[k]) < TIGHT, k
    empty = np.bincount(g["u"], minlength=g["n_users"]) == 0
    assert empty.any() and not np.allclose(g["E_theta"][empty], g["a0"] / g["b0"]), "fixture must cover the empty-row quirk"
    pred = O.poisson_ext_predict(g["val_u"], g["val_i"], st)
    assert rel_max(pred, g["val_pred"]) < TIGHT and (pred[:5] == 0).all()
    assert abs(O.rmse(g["val_x"], pred) - g["val_rmse"]) < TIGHT * g["val_rmse"]


def test_hpf_sweeps(golden):
    g = golden("hpf_cavi")
    cfg = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    st = O.hpf_sweeps(g["u"], g["i"], g["x"], g["K"], cfg, g["T"], g["seed"])
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi",
              "gamma_b_eta", "E_theta", "E_beta", "E_xi", "E_eta"):
        assert rel_max(st[k], g[k]) < TIGHT, k
    assert st["gamma_a_xi"] == pytest.approx(g["gamma_a_xi"], abs=0)
    pred = O.predict(g["val_u"], g["val_i"], st["E_theta"], st["E_beta"])
    assert rel_max(pred, g["val_pred"]) < TIGHT
    assert abs(O.rmse(g["val_x"], pred) - g["val_rmse"]) < 1e-12


@pytest.mark.parametrize("name,bias", [("gaussian_bias", True), ("gaussian_nobias", False)])
def test_gauss_sweeps(golden, name, bias):
    g = golden(name)
    st = O.gauss_sweeps(g["u"], g["i"], g["x"], g["K"], g["sigma2"], g["eta_theta2"], g["eta_beta2"],
                        g.get("eta_bias2", 1.0), g["T"], g["seed"], bias=bias)
    keys = ["m_theta", "V_theta", "m_beta", "V_beta"] + (["m_user_bias", "m_item_bias"] if bias else [])
    for k in keys:
        assert rel_max(st[k], g[k]) < 1e-10, k
    pred = O.predict(g["val_u"], g["val_i"], st["m_theta"], st["m_beta"],
                     st["m_user_bias"] if bias else None, st["m_item_bias"] if bias else None, g["global_mean"])
    assert rel_max(pred, g["val_pred"]) < 1e-10
    r, mm = O.gauss_eval(g["val_u"], g["val_i"], g["val_x"], st, g["global_mean"])
    assert abs(r - g["val_rmse"]) < 1e-10
    if bias:
        assert abs(mm - g["val_macro_mae"]) < 1e-10
        ok = (g["test_u"] < g["n_users"]) & (g["test_i"] < g["n_items"])
        lpl = O.gauss_lpl(g["test_u"][ok], g["test_i"][ok], g["test_x"][ok], st["m_theta"], st["m_beta"], g["sigma2"])
        assert abs(lpl - g["test_lpl"]) < 1e-9 * abs(g["test_lpl"])


@pytest.mark.parametrize("name,bias", [("gaussian_bias", True), ("gaussian_nobias", False)])
def test_gauss_sweeps_c_port(golden, name, bias):
    """oracle/pmf_oracle.c::orc_gauss_sweeps (Gauss-Jordan inverse instead of LAPACK) against the reference's outputs."""
    from oracle import c_oracle as CO
    g = golden(name)
    init = O.gauss_init(g["n_users"], g["n_items"], g["K"], g["seed"])
    st = CO.gauss_sweeps(g["u"], g["i"], g["x"], g["n_users"], g["n_items"], g["K"], g["sigma2"], g["eta_theta2"],
                         g["eta_beta2"], g.get("eta_bias2", 1.0), g["T"], init, bias=bias)
    keys = ["m_theta", "V_theta", "m_beta", "V_beta"] + (["m_user_bias", "m_item_bias"] if bias else [])
    for k in keys:
        assert rel_max(st[k], g[k]) < 1e-10, k


def test_hpf_map_loss_and_grads(golden):
    g = golden("hpf_pytorch")
    cfg = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    P = {k: g["init_" + k] for k in ("theta", "beta", "xi", "eta")}
    B = g["batch"]
    us = 1.0 / (g["user_counts"].astype(np.float32) + np.float32(1e-6))
    its = 1.0 / (g["item_counts"].astype(np.float32) + np.float32(1e-6))
    loss, G = O.hpf_map_loss_grads(P, g["u"][:B], g["i"][:B], g["x"][:B], us, its, cfg)
    assert abs(loss - g["first_loss"]) < 2e-5 * abs(g["first_loss"])   # torch side is fp32
    for k in G:
        assert rel_max(G[k], g["grad0_" + k]) < 5e-6, k


def test_hpf_map_training_replay(golden):
    """Full a6+a7 chain: torch RNG replay of the DataLoader shuffle + dense Adam, fp32."""
    torch = pytest.importorskip("torch")
    g = golden("hpf_pytorch")
    cfg = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    torch.manual_seed(g["torch_seed"])
    N, M, K = g["n_users"], g["n_items"], g["K"]
    # parameter creation order theta, beta, xi, eta (hpf_pytorch.py:39-48)
    P = {"theta": (torch.randn(N, K) * 0.1).numpy().astype(np.float64),
         "beta": (torch.randn(M, K) * 0.1).numpy().astype(np.float64),
         "xi": (torch.randn(N) * 0.1).numpy().astype(np.float64),
         "eta": (torch.randn(M) * 0.1).numpy().astype(np.float64)}
    for k in P:
        assert np.array_equal(P[k].astype(np.float32), g["init_" + k]), k
    us = (1.0 / (torch.tensor(g["user_counts"], dtype=torch.float32) + 1e-6)).numpy()
    its = (1.0 / (torch.tensor(g["item_counts"], dtype=torch.float32) + 1e-6)).numpy()
    Mo = {k: np.zeros_like(v) for k, v in P.items()}
    Vo = {k: np.zeros_like(v) for k, v in P.items()}
    n, B, step = len(g["u"]), g["batch"], 0
    for ep in range(g["epochs"]):
        # DataLoader.__iter__ draws _base_seed, then RandomSampler draws its own seed
        # (torch/utils/data/dataloader.py, sampler.py) -- both from the global CPU generator.
        _base_seed = int(torch.empty((), dtype=torch.int64).random_().item())
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        gen = torch.Generator(); gen.manual_seed(seed)
        perm = torch.randperm(n, generator=gen).numpy()
        if ep == 0:
            assert np.array_equal(g["u"][perm], g["epoch0_users"])
        tot = 0.0
        for s in range(0, n, B):
            idx = perm[s:s + B]
            step += 1
            loss, G = O.hpf_map_loss_grads(P, g["u"][idx], g["i"][idx], g["x"][idx], us, its, cfg)
            O.adam_dense_step(P, G, Mo, Vo, step, g["lr"])
            tot += loss
        assert abs(tot - g["epoch_loss"][ep]) < 1e-4 * abs(g["epoch_loss"][ep])
    for k in P:
        assert rel_max(P[k], g["final_" + k]) < 2e-4, k   # fp64 restatement vs fp32 torch Adam


# ---- the C restatement (oracle/pmf_oracle.c) against the same reference-generated vectors ----
def test_c_oracle_grouping(golden):
    from oracle import c_oracle as CO
    g = golden("poisson")
    rp, perm = CO.group(g["u"], g["n_users"])
    assert np.array_equal(perm, g["user_perm"]) and np.array_equal(np.diff(rp), g["user_counts"])
    rp, perm = CO.group(g["i"], g["n_items"])
    assert np.array_equal(perm, g["item_perm"]) and np.array_equal(np.diff(rp), g["item_counts"])


def test_c_oracle_poisson(golden):
    from oracle import c_oracle as CO
    g = golden("poisson")
    init = O.poisson_init(g["n_users"], g["n_items"], g["K"], g["a0"], g["b0"], g["seed"])
    st = CO.poisson_sweeps(g["u"], g["i"], g["x"], g["n_users"], g["n_items"], g["K"], g["a0"], g["b0"], g["T"],
                           init["E_theta"], init["E_beta"], threads=2)
    for k in ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta"):
        assert rel_max(st[k], g[k]) < TIGHT, k
    assert rel_max(CO.predict(g["val_u"], g["val_i"], st["E_theta"], st["E_beta"]), g["val_pred"]) < TIGHT


def test_c_oracle_hpf(golden):
    from oracle import c_oracle as CO
    g = golden("hpf_cavi")
    cfg = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    init = O.hpf_init(g["n_users"], g["n_items"], g["K"], cfg, g["seed"])
    st = CO.hpf_sweeps(g["u"], g["i"], g["x"], g["n_users"], g["n_items"], g["K"], cfg, g["T"], init, threads=2)
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
              "E_theta", "E_beta", "E_xi", "E_eta"):
        assert rel_max(st[k], g[k]) < TIGHT, k


def test_poisson_ext_scalar_rate_identity(golden):
    """The kernel never makes the reference's second in-row pass (poisson_mf_extended_cavi.py:160-164): by linearity
    b_phi = b0 + sum_t psi_t (beta_t . theta_new) = b0 + theta_new . (b_theta - b0).  Checked here on the reference's own
    outputs, in float64, for every row with observations (both sides)."""
    g = golden("poisson_ext")
    b0 = g["b0"]
    for E, b_vec, b_sc, ids, n in ((g["E_theta"], g["b_theta"], g["b_phi"], g["u"], g["n_users"]),
                                   (g["E_beta"], g["b_beta"], g["b_psi"], g["i"], g["n_items"])):
        seen = np.bincount(ids, minlength=n) > 0
        lhs = b_sc[seen]
        rhs = b0 + np.sum(E[seen] * (b_vec[seen] - b0), axis=1)
        assert rel_max(rhs, lhs) < 1e-12
