"""A few sweeps of BASELINE config C5 (or another CAVI workload) with the default tiling -- target for ncu captures."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from prob_matrix_factorization_b200 import synth
from prob_matrix_factorization_b200._engine import GammaEngine, row_stride
from prob_matrix_factorization_b200.ratings import DeviceRatings

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
n_sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w, (u, i, x) = synth.workload_ratings(name)
x = x + np.float32(1.0)
K = w.n_factors
dev = torch.device("cuda", 0)
dr = DeviceRatings(u, i, x, w.n_users, w.n_items, dev, row_bytes=4 * row_stride(K))
hyper = {"user_shape": 5.0 + K * 0.3, "user_rate_prior": 5.0, "item_shape": 5.0 + K * 0.3, "item_rate_prior": 5.0}
eng = GammaEngine(dr, K, 0.3, 0.3, None, None, hyper=hyper)
rng = np.random.default_rng(1)
eng.load_means(rng.random((w.n_users, K), dtype=np.float32) + 0.05, rng.random((w.n_items, K), dtype=np.float32) + 0.05,
               np.full(w.n_users, 1.3, np.float32), np.full(w.n_items, 0.9, np.float32))
torch.cuda.synchronize()
print(f"tiles user_pass={len(dr.user_tiles)} item_pass={len(dr.item_tiles)} launches/sweep={eng.launches_per_sweep}", flush=True)
for _ in range(n_sweeps):
    eng.sweep(False)
torch.cuda.synchronize()
