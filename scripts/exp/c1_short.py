"""A few Gaussian sweeps at C1 (for ncu launch lists)."""
import os, sys
import numpy as np, pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from prob_matrix_factorization_b200 import synth
from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
w, (u, i, x) = synth.workload_ratings("c1")
mean = float(x.mean())
df = pd.DataFrame({"u": u.astype(np.int64), "i": i.astype(np.int64), "rating": x.astype(np.float64) - mean})
hp = dict(sigma2=0.5, eta_theta2=0.1, eta_beta2=0.1, eta_bias2=0.1)
m = GaussianMFCAVI(GaussianMFCAVIConfig(n_factors=w.n_factors, max_iter=int(os.environ.get("C1_SWEEPS", 4)), tol=1e-3, verbose=False, **hp)).fit(df, global_mean=mean)
torch.cuda.synchronize()
e = m._engine
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    e.sweep(hp["sigma2"], hp["eta_theta2"], hp["eta_beta2"], hp["eta_bias2"])
e1.record(); torch.cuda.synchronize()
print(f"C1 sweep: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
