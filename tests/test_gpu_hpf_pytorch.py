"""a6/a7 on the GPU: the HPF_PyTorch drop-in against golden outputs of the reference module trained with the
scripts' own loop (torch.optim.Adam + DataLoader(shuffle=True), torch.manual_seed fixed by the harness)."""
import numpy as np
import pytest
import torch

from conftest import rel_max

pytestmark = pytest.mark.gpu
P = ("theta", "beta", "xi", "eta")


def build(g):
    from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config
    hp = {k: g[k] for k in ("a", "a_prime", "b_prime", "c", "c_prime", "d_prime")}
    cfg = HPF_PyTorch_Config(n_factors=g["K"], lr=g["lr"], batch_size=g["batch"], epochs=g["epochs"], verbose=False, **hp)
    torch.manual_seed(g["torch_seed"])
    return HPF_PyTorch(g["n_users"], g["n_items"], g["user_counts"], g["item_counts"], cfg)


def test_init_is_bit_identical(golden):
    g = golden("hpf_pytorch")
    m = build(g)
    for k in P:
        p = getattr(m, k + "_uncons")
        assert p.is_cuda and isinstance(p, torch.nn.Parameter)
        assert np.array_equal(p.detach().cpu().numpy(), g["init_" + k]), k
    assert [n for n, _ in m.named_parameters()] == ["theta_uncons", "beta_uncons", "xi_uncons", "eta_uncons"]


def test_loss_and_backward_first_batch(golden):
    g = golden("hpf_pytorch")
    m = build(g)
    B = g["batch"]
    loss = m.loss(torch.LongTensor(g["u"][:B]), torch.LongTensor(g["i"][:B]), torch.FloatTensor(g["x"][:B]))
    assert loss.dim() == 0 and loss.dtype == torch.float32
    loss.backward()
    assert abs(loss.item() - g["first_loss"]) < 1e-5 * abs(g["first_loss"])
    for k in P:
        assert rel_max(getattr(m, k + "_uncons").grad.cpu().numpy(), g["grad0_" + k]) < 1e-5, k


def test_script_loop_unchanged(golden):
    """compare_models.py:288-313 verbatim on top of the drop-in."""
    g = golden("hpf_pytorch")
    m = build(g)
    opt = torch.optim.Adam(m.parameters(), lr=g["lr"])

    class DS(torch.utils.data.Dataset):
        def __init__(s):
            s.u = torch.LongTensor(g["u"]); s.i = torch.LongTensor(g["i"]); s.r = torch.FloatTensor(g["x"])
        def __len__(s): return len(s.r)
        def __getitem__(s, idx): return s.u[idx], s.i[idx], s.r[idx]

    loader = torch.utils.data.DataLoader(DS(), batch_size=g["batch"], shuffle=True)
    for ep in range(g["epochs"]):
        m.train(); tot = 0.0; first = []
        for users, items, ratings in loader:
            if ep == 0:
                first.append(users.numpy().copy())
            opt.zero_grad()
            loss = m.loss(users, items, ratings)
            loss.backward(); opt.step(); tot += loss.item()
        if ep == 0:
            assert np.array_equal(np.concatenate(first), g["epoch0_users"])     # same shuffle as the golden run
        assert abs(tot - g["epoch_loss"][ep]) < 1e-5 * abs(g["epoch_loss"][ep])
    m.eval()
    for k in P:
        assert rel_max(getattr(m, k + "_uncons").detach().cpu().numpy(), g["final_" + k]) < 2e-5, k
    vu = np.minimum(g["val_u"], g["n_users"] - 1); vi = np.minimum(g["val_i"], g["n_items"] - 1)
    pred = m.predict(vu, vi)
    assert pred.dtype == np.float32 and rel_max(pred, g["val_pred"]) < 2e-5
    assert rel_max(m.theta.detach().cpu().numpy(), np.logaddexp(0, g["final_theta"])) < 2e-5


@pytest.mark.parametrize("lazy", [False, True])
def test_fit_epochs_replays_the_loader(golden, lazy):
    g = golden("hpf_pytorch")
    m = build(g)
    losses = m.fit_epochs(g["u"], g["i"], g["x"], epochs=g["epochs"], batch_size=g["batch"], lr=g["lr"], lazy=lazy)
    for ep in range(g["epochs"]):
        assert abs(losses[ep] - g["epoch_loss"][ep]) < 1e-5 * abs(g["epoch_loss"][ep])
    for k in P:
        assert rel_max(getattr(m, k + "_uncons").detach().cpu().numpy(), g["final_" + k]) < 2e-5, k


def test_wide_factors_and_out_of_range_ids():
    from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config
    from oracle import pmf_oracle as O
    rng = np.random.default_rng(3)
    N, M, K, B = 50, 40, 100, 300
    cfg = HPF_PyTorch_Config(n_factors=K, a=0.3, c=0.3)
    torch.manual_seed(5)
    m = HPF_PyTorch(N, M, rng.integers(1, 9, N), rng.integers(1, 9, M), cfg)
    u = rng.integers(0, N, B); i = rng.integers(0, M, B); r = rng.integers(1, 7, B).astype(np.float32)
    loss = m.loss(torch.LongTensor(u), torch.LongTensor(i), torch.FloatTensor(r)); loss.backward()
    Pm = {k: getattr(m, k + "_uncons").detach().cpu().numpy() for k in P}
    hp = dict(a=0.3, a_prime=1.0, b_prime=1.0, c=0.3, c_prime=1.0, d_prime=1.0)
    ref_loss, G = O.hpf_map_loss_grads(Pm, u, i, r, m.user_scale.cpu().numpy(), m.item_scale.cpu().numpy(), hp)
    assert abs(loss.item() - ref_loss) < 1e-5 * abs(ref_loss)
    for k in P:
        assert rel_max(getattr(m, k + "_uncons").grad.cpu().numpy(), G[k]) < 1e-5, k
    m.loss(torch.LongTensor([N + 3]), torch.LongTensor([0]), torch.FloatTensor([1.0]))
    with pytest.raises(IndexError):
        m.check_ids()


def test_lazy_adam_equals_dense_adam():
    """Touch-only Adam replays zero-gradient steps exactly: same parameters and moments as the dense kernel,
    also across two fit_epochs calls and a switch between the two modes (tolerance covers only the float
    atomics that combine duplicate ids inside a batch)."""
    from prob_matrix_factorization_b200 import synth
    from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config
    N, M, nnz, K = 3000, 2500, 20_000, 100
    u, i, x = synth.make_ratings(N, M, nnz, seed=77)
    uc = np.bincount(u, minlength=N); ic = np.bincount(i, minlength=M)
    cfg = HPF_PyTorch_Config(n_factors=K, a=0.3, c=0.3, lr=0.01)
    out = {}
    for mode in ("dense", "lazy", "mixed"):
        torch.manual_seed(11)
        m = HPF_PyTorch(N, M, uc, ic, cfg)
        l1 = m.fit_epochs(u, i, x + 1.0, epochs=2, batch_size=512, lr=0.01, lazy=mode != "dense")
        l2 = m.fit_epochs(u, i, x + 1.0, epochs=1, batch_size=512, lr=0.01, lazy=mode == "lazy")
        out[mode] = ([getattr(m, k + "_uncons").detach().cpu().numpy() for k in P], [t.cpu().numpy() for t in m._adam["m"]],
                     [t.cpu().numpy() for t in m._adam["v"]], l1 + l2)
    for mode in ("lazy", "mixed"):
        for a, b in zip(out["dense"][:3], out[mode][:3]):
            for x_, y_ in zip(a, b):
                assert rel_max(y_, x_) < 1e-5, mode
        assert np.allclose(out["dense"][3], out[mode][3], rtol=1e-6)
