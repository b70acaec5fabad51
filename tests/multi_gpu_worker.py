"""Worker for tests/test_gpu_multi.py: launched by torchrun, one rank per GPU (NCCL)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import c_oracle as CO  # noqa: E402
from prob_matrix_factorization_b200 import synth  # noqa: E402
from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config  # noqa: E402
from prob_matrix_factorization_b200.parallel import init_process_group  # noqa: E402
from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig  # noqa: E402


def rel_max(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    os.environ["PMF_EXCHANGE"] = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    rank, world, local = init_process_group()
    dev = torch.device("cuda", local)
    N, M, nnz, K, T = 30_000, 12_000, 400_000, 24, 10
    u, i, x = synth.make_ratings(N, M, nnz, seed=99)
    x = x + 1.0
    hp = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)
    m = HPF_CAVI(HPF_CAVI_Config(n_factors=K, max_iter=T, tol=None, random_state=42, verbose=False, **hp),
                 device=dev, shard=(rank, world), seg_len=64)
    m.n_users, m.n_items = N, M
    init = m._initial_state()
    m.fit_arrays(u, i, x, init)
    ref = CO.hpf_sweeps(u, i, x, N, M, K, hp, T, init)
    worst = 0.0
    for k in ("gamma_a_theta", "gamma_b_theta", "gamma_a_beta", "gamma_b_beta", "gamma_b_xi", "gamma_b_eta",
              "E_theta", "E_beta", "E_xi", "E_eta"):
        worst = max(worst, rel_max(getattr(m, k), ref[k]))
    assert m._engine.exchange == ("nccl" if os.environ["PMF_EXCHANGE"] == "nccl" else "closed")
    # every rank must hold the same replicated tables bit for bit
    h = torch.stack([m._engine.E_theta.double().sum(), m._engine.E_beta.double().sum()])
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    same = all(torch.equal(hs[0], t) for t in hs)
    p = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=10, a0=0.1, b0=0.5, max_iter=5, tol=None, verbose=False),
                      device=dev, shard=(rank, world))
    p.n_users, p.n_items = N, M
    pin = p._initial_state()
    p.fit_arrays(u, i, x, pin)
    pref = CO.poisson_sweeps(u, i, x, N, M, 10, 0.1, 0.5, 5, pin["E_theta"], pin["E_beta"])
    worst_p = max(rel_max(p.E_theta, pref["E_theta"]), rel_max(p.a_beta, pref["a_beta"]), rel_max(p.b_theta, pref["b_theta"]))
    ok = worst < 1e-5 and worst_p < 1e-5 and same
    print(f"rank {rank}/{world}: hpf rel err {worst:.2e}, poisson rel err {worst_p:.2e}, replicas identical {same} -> "
          f"{'MULTI_GPU_OK' if ok else 'MULTI_GPU_FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
