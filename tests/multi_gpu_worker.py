"""Worker for tests/test_gpu_multi.py: launched by torchrun, one rank per GPU (NCCL).

Ratings sharded by user range (each rank uploads 1/world of the list, one all-to-all routes every rating to its user's
owner), local user pass, item pass + cross-rank combine of the row sums.  Checks, on every rank:
  A  HPF, 10 sweeps, default configuration                      vs the oracle's C port, <= 1e-5
  B  HPF with forced tiles (2 item tiles / 3 user tiles) and 3 item chunks (pipelined combine)   likewise
  C  Poisson MF, 5 sweeps                                        likewise
  D  early stopping on a validation frame: statistics are summed over the ranks, every rank stops at the same sweep,
     and the last validation RMSE equals the oracle's at that sweep
  E  ELBO of the sharded state (a11, parity unpinned)            vs the oracle's restatement
and that the replicated tables are bit-identical on all ranks.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import c_oracle as CO  # noqa: E402
from oracle import pmf_oracle as O  # noqa: E402
from prob_matrix_factorization_b200 import synth  # noqa: E402
from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config  # noqa: E402
from prob_matrix_factorization_b200.parallel import init_process_group  # noqa: E402
from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig  # noqa: E402

HP = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
HPF_TABLES = ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
              "E_theta", "E_beta", "E_xi", "E_eta")


def rel_max(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def replicas_identical(m, world):
    e = m._engine
    h = torch.stack([e.E_theta.double().sum(), e.E_beta.double().sum(), e.shp_beta.double().sum(), e.rte_theta.double().sum()])
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    return all(torch.equal(hs[0], t) for t in hs)


def main():
    exchange = sys.argv[1] if len(sys.argv) > 1 else "mc"
    os.environ["PMF_EXCHANGE"] = exchange
    rank, world, local = init_process_group()
    dev = torch.device("cuda", local)
    N, M, nnz, K, T = 30_000, 12_000, 400_000, 24, 10
    u, i, x = synth.make_ratings(N, M, nnz, seed=99)
    x = x + 1.0
    report = {}

    def hpf(T, tol=None, K=K, **kw):
        m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=tol, random_state=42, verbose=False, **HP),
                     device=dev, shard=(rank, world), seg_len=64, **kw)
        return m

    # A: default configuration
    m = hpf(T)
    m.n_users, m.n_items = N, M
    init = m._initial_state()
    m.fit_arrays(u, i, x, init)
    ref = CO.hpf_sweeps(u, i, x, N, M, K, HP, T, init)
    report["A"] = max(rel_max(getattr(m, k), ref[k]) for k in HPF_TABLES)
    same = replicas_identical(m, world)
    want = "closed"
    ok_state = m._engine.exchange == want

    # B: forced tiles + pipelined chunks
    mb = hpf(T)
    mb.n_users, mb.n_items = N, M
    mb._ratings_kw = dict(user_pass_tiles=2, item_pass_tiles=3, item_chunks=3)
    mb.fit_arrays(u, i, x, init)
    report["B"] = max(rel_max(getattr(mb, k), ref[k]) for k in HPF_TABLES)
    same = same and replicas_identical(mb, world)

    # C: Poisson
    p = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=10, a0=0.1, b0=0.5, max_iter=5, tol=None, verbose=False),
                      device=dev, shard=(rank, world))
    p.n_users, p.n_items = N, M
    pin = p._initial_state()
    p.fit_arrays(u, i, x, pin)
    pref = CO.poisson_sweeps(u, i, x, N, M, 10, 0.1, 0.5, 5, pin["E_theta"], pin["E_beta"])
    report["C"] = max(rel_max(getattr(p, k), pref[k]) for k in ("a_theta", "b_theta", "a_beta", "b_beta", "E_theta", "E_beta"))

    # D: early stopping with a validation frame (+ unseen ids, which predict 0 and still count)
    vu, vi, vx = synth.make_ratings(N, M, 40_000, seed=7)
    vu = np.concatenate([vu, [N + 5, 3]]).astype(np.int64); vi = np.concatenate([vi, [2, M + 9]]).astype(np.int64)
    vx = np.concatenate([vx + 1.0, [4.0, 2.0]])
    md = hpf(30, tol=2e-3)
    md.n_users, md.n_items = N, M
    md.fit_arrays(u, i, x, init, val=(vu, vi, vx))
    iters = torch.tensor([md.n_iter_], device=dev)
    all_iters = [torch.zeros_like(iters) for _ in range(world)]
    dist.all_gather(all_iters, iters)
    refd = CO.hpf_sweeps(u, i, x, N, M, K, HP, md.n_iter_, init)
    ok_ids = (vu < N) & (vi < M)
    pred = np.where(ok_ids, CO.predict(np.minimum(vu, N - 1), np.minimum(vi, M - 1), refd["E_theta"], refd["E_beta"]), 0.0)
    rmse_ref = float(np.sqrt(np.mean((vx - pred) ** 2)))
    report["D"] = abs(md.val_rmse_history_[-1] - rmse_ref) / rmse_ref
    same_stop = all(int(t.item()) == md.n_iter_ for t in all_iters) and 1 < md.n_iter_ < 30
    report["D_E_theta"] = rel_max(md.E_theta, refd["E_theta"])

    # E: ELBO of a sharded state (small problem: the oracle's ELBO is a NumPy restatement)
    Ns, Ms = 400, 300
    us, is_, xs = synth.make_ratings(Ns, Ms, 8000, 21)
    xs = xs + 1.0
    me = hpf(4, K=6, track_elbo=True)
    me.n_users, me.n_items = Ns, Ms
    me.fit_arrays(us, is_, xs)
    hist = [O.hpf_elbo(us, is_, xs, O.hpf_sweeps(us, is_, xs, 6, HP, t, 42, Ns, Ms), HP) for t in range(1, 5)]
    report["E"] = rel_max(np.array(me.elbo_history_), np.array(hist))

    # F: Gaussian MF (with and without biases) sharded: item statistics all-reduced over NCCL; early stopping on a sharded
    #    validation frame takes the same decision on every rank
    from prob_matrix_factorization_b200 import gaussian_mf_cavi as NB
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    import pandas as pd
    Ng, Mg, Kg, Tg = 3000, 1800, 10, 6
    ug, ig, xg = synth.make_ratings(Ng, Mg, 40_000, seed=5)
    mean = float(xg.mean())
    xc = xg.astype(np.float64) - mean
    frame = pd.DataFrame({"u": ug.astype(np.int64), "i": ig.astype(np.int64), "rating": xc})
    ghp = dict(sigma2=0.5, eta_theta2=0.1, eta_beta2=0.1)
    for bias in (True, False):
        if bias:
            mg = GaussianMFCAVI(GaussianMFCAVIConfig(n_factors=Kg, eta_bias2=0.1, max_iter=Tg, random_state=42, verbose=False, **ghp),
                                device=dev, shard=(rank, world))
        else:
            mg = NB.GaussianMFCAVI(NB.GaussianMFCAVIConfig(n_factors=Kg, max_iter=Tg, random_state=42, verbose=False, **ghp),
                                   device=dev, shard=(rank, world))
        mg.fit(frame, global_mean=mean)
        gref = CO.gauss_sweeps(ug, ig, xc, Ng, Mg, Kg, 0.5, 0.1, 0.1, 0.1, Tg, O.gauss_init(Ng, Mg, Kg, 42), bias=bias)
        keys = ["m_theta", "V_theta", "m_beta", "V_beta"] + (["m_user_bias", "m_item_bias"] if bias else [])
        report["F_bias" if bias else "F_nobias"] = max(rel_max(getattr(mg, k), gref[k]) for k in keys)

    worst = max(report.values())
    ok = worst < 1e-5 and same and same_stop and ok_state
    print(f"rank {rank}/{world} [{exchange}]: " + " ".join(f"{k}={v:.2e}" for k, v in report.items())
          + f" replicas identical {same}, same stop sweep {same_stop} (n_iter {md.n_iter_}) -> "
          f"{'MULTI_GPU_OK' if ok else 'MULTI_GPU_FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
