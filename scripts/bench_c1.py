"""BASELINE config C1 timing (gaussian_mf with biases, K=10, 20k users x 10k recipes x 200k ratings): CAVI sweeps
through the drop-in on the GPU (CUDA events around engine.sweep, state resident) and end to end through ``fit`` from a
DataFrame, next to the oracle's reference-style NumPy row loops on the host (1 sweep, single core -- the reference's
own loop is the same per-row Python + np.linalg.inv).  Not part of the product; numbers go to profiles/README.md."""
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import pmf_oracle as O  # noqa: E402  (CPU baseline leg only)
from prob_matrix_factorization_b200 import synth  # noqa: E402
from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig  # noqa: E402

w, (u, i, x) = synth.workload_ratings("c1")
mean = float(x.mean())
xc = (x - mean).astype(np.float64)                      # compare_models.py:54-58: ratings are centred by the caller
K, nnz, T = w.n_factors, w.nnz, 20
hp = dict(sigma2=0.5, eta_theta2=0.1, eta_beta2=0.1, eta_bias2=0.1)      # best_hyperparams.txt:3
df = pd.DataFrame({"u": u.astype(np.int64), "i": i.astype(np.int64), "rating": xc})
cfg = GaussianMFCAVIConfig(n_factors=K, max_iter=T, tol=1e-3, random_state=42, verbose=False, **hp)
m = GaussianMFCAVI(cfg).fit(df, global_mean=mean)       # warm-up: context, allocator, first launches
torch.cuda.synchronize()
t = time.perf_counter()
m = GaussianMFCAVI(cfg).fit(df, global_mean=mean)
m_theta = m.m_theta                                     # D2H of the factors, as train_gaussian_full.py:77 reads them
dt = time.perf_counter() - t
print(f"C1 fit end to end (DataFrame in, {T} sweeps, m_theta out): {dt * 1e3:.1f} ms -> {nnz * T / dt:.3e} nnz*iters/s", flush=True)
eng = m._engine
for _ in range(3):
    eng.sweep(hp["sigma2"], hp["eta_theta2"], hp["eta_beta2"], hp["eta_bias2"])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
R = 50
e0.record()
for _ in range(R):
    eng.sweep(hp["sigma2"], hp["eta_theta2"], hp["eta_beta2"], hp["eta_bias2"])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
alg = 2 * (nnz * (4 * (K * (K + 1) // 2 + K) + 12)) + (w.n_users + w.n_items) * 4 * (K * K + K) + 2 * (nnz * (4 * K + 12)) + (w.n_users + w.n_items) * 12
print(f"C1 device sweep: {ms * 1e3:.1f} us -> {nnz / ms * 1e3:.3e} nnz*iters/s; algorithmic bytes {alg / 1e6:.0f} MB/sweep -> "
      f"{alg / ms / 1e6:.0f} GB/s", flush=True)
t = time.perf_counter()
O.gauss_sweeps(u, i, xc, K, hp["sigma2"], hp["eta_theta2"], hp["eta_beta2"], hp["eta_bias2"], 1, 42)
dt = time.perf_counter() - t
print(f"C1 CPU reference-style row loops (oracle, 1 core, incl. grouping): 1 sweep in {dt:.2f} s -> {nnz / dt:.3e} nnz*iters/s", flush=True)
