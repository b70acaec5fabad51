"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes.

Run in the build container only (the reference is mounted read-only there):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``src.models.*`` / ``src.evaluation.metrics`` from /root/reference, fits
each model on small seeded synthetic rating lists and stores inputs + outputs.
The GPU box has no /root/reference; tests there read only the committed .npz.
Library versions used are recorded in ``versions.json`` (the reference's own
requirements.txt pins are older).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PMF_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

import pandas as pd  # noqa: E402
import torch  # noqa: E402

from prob_matrix_factorization_b200 import synth  # noqa: E402
from src.evaluation import metrics as ref_metrics  # noqa: E402
from src.models.gaussian_mf_cavi import GaussianMFCAVI as RefGaussNoBias  # noqa: E402
from src.models.gaussian_mf_cavi import GaussianMFCAVIConfig as RefGaussNoBiasCfg  # noqa: E402
from src.models.gaussian_mf_cavi_bias import GaussianMFCAVI as RefGauss  # noqa: E402
from src.models.gaussian_mf_cavi_bias import GaussianMFCAVIConfig as RefGaussCfg  # noqa: E402
from src.models.hpf_cavi import HPF_CAVI as RefHPF  # noqa: E402
from src.models.hpf_cavi import HPF_CAVI_Config as RefHPFCfg  # noqa: E402
from src.models.hpf_pytorch import HPF_PyTorch as RefHPFTorch  # noqa: E402
from src.models.hpf_pytorch import HPF_PyTorch_Config as RefHPFTorchCfg  # noqa: E402
from src.models.poisson_mf_cavi import PoissonMFCAVI as RefPoisson  # noqa: E402
from src.models.poisson_mf_cavi import PoissonMFCAVIConfig as RefPoissonCfg  # noqa: E402
from src.models.poisson_mf_extended_cavi import PoissonMFExtendedCAVI as RefPoissonExt  # noqa: E402
from src.models.poisson_mf_extended_cavi import PoissonMFExtendedCAVIConfig as RefPoissonExtCfg  # noqa: E402

N_USERS, N_ITEMS, NNZ, SEED = 300, 400, 1500, 777


def frames(shift=0.0):
    (u, i, x), (vu, vi, vx), (tu, ti, tx) = synth.make_splits(N_USERS, N_ITEMS, NNZ, SEED)
    # a few validation rows with ids the model has never seen (predict -> 0, still scored)
    vu = vu.copy(); vi = vi.copy()
    vu[:3] = N_USERS + np.arange(3)
    vi[3:5] = N_ITEMS + 7
    f = lambda a, b, c: pd.DataFrame({"u": a.astype(np.int64), "i": b.astype(np.int64),
                                      "rating": c.astype(np.float64) + shift})
    return f(u, i, x), f(vu, vi, vx), f(tu, ti, tx)


def iterations_run(fn):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn()
    return sum(1 for ln in buf.getvalue().splitlines() if "iteration " in ln), buf.getvalue()


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", ()) for k, v in arrays.items()})


def base_inputs(tr, va, te):
    return dict(u=tr["u"].to_numpy(), i=tr["i"].to_numpy(), x=tr["rating"].to_numpy(),
                val_u=va["u"].to_numpy(), val_i=va["i"].to_numpy(), val_x=va["rating"].to_numpy(),
                test_u=te["u"].to_numpy(), test_i=te["i"].to_numpy(), test_x=te["rating"].to_numpy())


def golden_poisson():
    tr, va, te = frames()
    K, T = 7, 6
    cfg = RefPoissonCfg(n_factors=K, a0=0.1, b0=0.5, max_iter=T, tol=None, random_state=42, verbose=False)
    m = RefPoisson(cfg).fit(tr)
    # grouping lists straight from the reference helper
    uo, io_ = m._build_index_lists(tr["u"].to_numpy(), tr["i"].to_numpy(), m.n_users, m.n_items)
    out = base_inputs(tr, va, te)
    out.update(K=K, T=T, a0=0.1, b0=0.5, seed=42, n_users=m.n_users, n_items=m.n_items,
               a_theta=m.a_theta, b_theta=m.b_theta, a_beta=m.a_beta, b_beta=m.b_beta,
               E_theta=m.E_theta, E_beta=m.E_beta,
               user_perm=np.concatenate(uo), user_counts=np.array([len(v) for v in uo]),
               item_perm=np.concatenate(io_), item_counts=np.array([len(v) for v in io_]),
               val_pred=m.predict(va["u"].to_numpy(), va["i"].to_numpy()),
               val_rmse=m.evaluate_rmse(va), val_macro_mae=m.evaluate_macro_mae(va),
               test_rmse=m.evaluate_rmse(te),
               test_lpl=ref_metrics.PoissonLogPredictiveLikelihood(te, m.E_theta, m.E_beta))
    # early stopping semantics (:213): stops as soon as improvement < tol (even negative)
    cfg2 = RefPoissonCfg(n_factors=K, a0=0.1, b0=0.5, max_iter=40, tol=2e-3, random_state=42, verbose=True)
    m2 = RefPoisson(cfg2)
    n_it, _ = iterations_run(lambda: m2.fit(tr, va))
    out.update(es_tol=2e-3, es_max_iter=40, es_iterations=n_it, es_E_theta=m2.E_theta, es_val_rmse=m2.evaluate_rmse(va))
    save("poisson", **out)


def golden_poisson_ext():
    tr, va, te = frames()
    K, T = 7, 6
    names = ("a_theta", "b_theta", "a_beta", "b_beta", "a_phi", "b_phi", "a_psi", "b_psi", "E_theta", "E_beta", "E_phi", "E_psi")
    cfg = RefPoissonExtCfg(n_factors=K, a0=0.3, b0=1.0, max_iter=T, tol=None, random_state=42, verbose=False)
    m = RefPoissonExt(cfg).fit(tr)
    out = base_inputs(tr, va, te)
    out.update(K=K, T=T, a0=0.3, b0=1.0, seed=42, n_users=m.n_users, n_items=m.n_items,
               val_pred=m.predict(va["u"].to_numpy(), va["i"].to_numpy()), val_rmse=m.evaluate_rmse(va),
               test_rmse=m.evaluate_rmse(te), **{k: getattr(m, k) for k in names})
    cfg2 = RefPoissonExtCfg(n_factors=K, a0=0.3, b0=1.0, max_iter=40, tol=2e-3, random_state=42, verbose=True)
    m2 = RefPoissonExt(cfg2)
    n_it, _ = iterations_run(lambda: m2.fit(tr, va))
    out.update(es_tol=2e-3, es_max_iter=40, es_iterations=n_it, es_E_theta=m2.E_theta, es_E_phi=m2.E_phi,
               es_val_rmse=m2.evaluate_rmse(va))
    save("poisson_ext", **out)


def golden_hpf():
    tr, va, te = frames(shift=1.0)  # compare_models.py:180-185
    K, T = 6, 6
    hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
    cfg = RefHPFCfg(n_factors=K, max_iter=T, tol=None, random_state=42, verbose=False, **hp)
    m = RefHPF(cfg).fit(tr)
    out = base_inputs(tr, va, te)
    out.update(K=K, T=T, seed=42, n_users=m.n_users, n_items=m.n_items, **hp,
               gamma_a_theta=m.gamma_a_theta, gamma_b_theta=m.gamma_b_theta,
               gamma_a_beta=m.gamma_a_beta, gamma_b_beta=m.gamma_b_beta,
               gamma_a_xi=m.gamma_a_xi, gamma_b_xi=m.gamma_b_xi,
               gamma_a_eta=m.gamma_a_eta, gamma_b_eta=m.gamma_b_eta,
               E_theta=m.E_theta, E_beta=m.E_beta, E_xi=m.E_xi, E_eta=m.E_eta,
               val_pred=m.predict(va["u"].to_numpy(), va["i"].to_numpy()),
               val_rmse=m.evaluate_rmse(va), val_macro_mae=m.evaluate_macro_mae(va),
               test_rmse=m.evaluate_rmse(te))
    cfg2 = RefHPFCfg(n_factors=K, max_iter=40, tol=2e-3, random_state=42, verbose=True, **hp)
    m2 = RefHPF(cfg2)
    n_it, _ = iterations_run(lambda: m2.fit(tr, va))
    out.update(es_tol=2e-3, es_max_iter=40, es_iterations=n_it, es_E_theta=m2.E_theta, es_val_rmse=m2.evaluate_rmse(va))
    save("hpf_cavi", **out)


def golden_gauss():
    tr, va, te = frames()
    mean = tr["rating"].mean()                              # compare_models.py:54-65
    for d in (tr, va, te):
        d["rating"] -= mean
    K, T = 5, 5
    hp = dict(sigma2=0.3, eta_theta2=0.5, eta_beta2=0.5)
    cfg = RefGaussCfg(n_factors=K, eta_bias2=1.0, max_iter=T, tol=1e-3, random_state=42, verbose=False, **hp)
    m = RefGauss(cfg).fit(tr, global_mean=mean)
    out = base_inputs(tr, va, te)
    out.update(K=K, T=T, seed=42, eta_bias2=1.0, global_mean=mean, n_users=m.n_users, n_items=m.n_items, **hp,
               m_theta=m.m_theta, V_theta=m.V_theta, m_beta=m.m_beta, V_beta=m.V_beta,
               m_user_bias=m.m_user_bias, m_item_bias=m.m_item_bias,
               val_pred=m.predict(va["u"].to_numpy(), va["i"].to_numpy(), mean),
               val_rmse=m.evaluate_rmse(va, mean), val_macro_mae=m.evaluate_macro_mae(va, mean),
               test_rmse=m.evaluate_rmse(te, mean),
               test_lpl=ref_metrics.GaussianLogPredictiveLikelihood(
                   te[(te.u < m.n_users) & (te.i < m.n_items)], m.m_theta, m.m_beta, hp["sigma2"]))
    # early stop (:279) needs 0 <= improvement < tol
    cfg2 = RefGaussCfg(n_factors=K, eta_bias2=1.0, max_iter=30, tol=5e-3, random_state=42, verbose=True, **hp)
    m2 = RefGauss(cfg2)
    n_it, _ = iterations_run(lambda: m2.fit(tr, va, global_mean=mean))
    out.update(es_tol=5e-3, es_max_iter=30, es_iterations=n_it, es_m_theta=m2.m_theta,
               es_val_rmse=m2.evaluate_rmse(va, mean))
    save("gaussian_bias", **out)

    cfg3 = RefGaussNoBiasCfg(n_factors=K, max_iter=T, tol=1e-3, random_state=42, verbose=False, **hp)
    m3 = RefGaussNoBias(cfg3).fit(tr, global_mean=mean)
    out3 = base_inputs(tr, va, te)
    out3.update(K=K, T=T, seed=42, n_users=m3.n_users, n_items=m3.n_items, **hp,
                m_theta=m3.m_theta, V_theta=m3.V_theta, m_beta=m3.m_beta, V_beta=m3.V_beta,
                global_mean=mean,
                val_pred=m3.predict(va["u"].to_numpy(), va["i"].to_numpy(), mean),
                val_rmse=m3.evaluate_rmse(va, mean))
    save("gaussian_nobias", **out3)


def golden_hpf_torch():
    tr, va, te = frames(shift=1.0)                          # compare_models.py:243-249
    n_users = int(max(tr["u"].max(), va["u"].max(), te["u"].max()) + 1)   # :251-252
    n_items = int(max(tr["i"].max(), va["i"].max(), te["i"].max()) + 1)
    user_counts = np.zeros(n_users); item_counts = np.zeros(n_items)
    uv, uc = np.unique(tr["u"], return_counts=True); user_counts[uv] = uc
    iv, ic = np.unique(tr["i"], return_counts=True); item_counts[iv] = ic
    K, EPOCHS, BATCH, LR, TSEED = 4, 3, 256, 0.01, 1234
    hp = dict(a=0.3, a_prime=1.0, b_prime=1.0, c=0.3, c_prime=1.0, d_prime=1.0)
    cfg = RefHPFTorchCfg(n_factors=K, lr=LR, batch_size=BATCH, epochs=EPOCHS, verbose=False, **hp)
    torch.manual_seed(TSEED)
    model = RefHPFTorch(n_users, n_items, user_counts, item_counts, cfg)
    init = {k: getattr(model, k + "_uncons").detach().numpy().copy() for k in ("theta", "beta", "xi", "eta")}
    opt = torch.optim.Adam(model.parameters(), lr=LR)

    class DS(torch.utils.data.Dataset):                     # compare_models.py:291-297
        def __init__(self, df):
            self.u = torch.LongTensor(df["u"].values); self.i = torch.LongTensor(df["i"].values)
            self.r = torch.FloatTensor(df["rating"].values)
        def __len__(self): return len(self.r)
        def __getitem__(self, idx): return self.u[idx], self.i[idx], self.r[idx]

    loader = torch.utils.data.DataLoader(DS(tr), batch_size=BATCH, shuffle=True)
    # loss/grad of the very first batch in file order, before any step (pure a6 check)
    fb_u = torch.LongTensor(tr["u"].values[:BATCH]); fb_i = torch.LongTensor(tr["i"].values[:BATCH])
    fb_r = torch.FloatTensor(tr["rating"].values[:BATCH])
    model.zero_grad()
    l0 = model.loss(fb_u, fb_i, fb_r); l0.backward()
    g0 = {k: getattr(model, k + "_uncons").grad.numpy().copy() for k in ("theta", "beta", "xi", "eta")}
    model.zero_grad()
    epoch_loss, batches = [], []
    for ep in range(EPOCHS):                                # compare_models.py:305-313
        model.train(); tot = 0.0
        for users, items, ratings in loader:
            if ep == 0:
                batches.append(users.numpy().copy())
            opt.zero_grad()
            loss = model.loss(users, items, ratings)
            loss.backward(); opt.step(); tot += loss.item()
        epoch_loss.append(tot)
    model.eval()
    final = {k: getattr(model, k + "_uncons").detach().numpy().copy() for k in ("theta", "beta", "xi", "eta")}
    out = base_inputs(tr, va, te)
    out.update(K=K, epochs=EPOCHS, batch=BATCH, lr=LR, torch_seed=TSEED, n_users=n_users, n_items=n_items,
               user_counts=user_counts, item_counts=item_counts, **hp,
               first_loss=float(l0.item()), epoch_loss=np.array(epoch_loss),
               epoch0_users=np.concatenate(batches),
               val_pred=model.predict(va["u"].values.clip(max=n_users - 1), va["i"].values.clip(max=n_items - 1)),
               **{"init_" + k: v for k, v in init.items()},
               **{"grad0_" + k: v for k, v in g0.items()},
               **{"final_" + k: v for k, v in final.items()})
    save("hpf_pytorch", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:            # python make_golden.py poisson_ext  -> only that fixture (versions.json untouched)
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_poisson()
    golden_poisson_ext()
    golden_hpf()
    golden_gauss()
    golden_hpf_torch()
    import scipy
    with open(os.path.join(HERE, "versions.json"), "w") as f:
        json.dump({"numpy": np.__version__, "pandas": pd.__version__, "torch": torch.__version__,
                   "scipy": scipy.__version__, "python": sys.version.split()[0],
                   "reference_root": REF, "sizes": [N_USERS, N_ITEMS, NNZ, SEED]}, f, indent=1)
