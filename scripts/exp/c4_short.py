"""Two lazy-Adam epochs at C4 (for ncu launch lists / quick timing)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from prob_matrix_factorization_b200 import synth
from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config

w, (u, i, x) = synth.workload_ratings("c4")
x = x + 1.0
N, M, K = w.n_users, w.n_items, w.n_factors
cfg = HPF_PyTorch_Config(n_factors=K, a=0.3, c=0.3, lr=5e-4)
torch.manual_seed(0)
m = HPF_PyTorch(N, M, np.bincount(u, minlength=N), np.bincount(i, minlength=M), cfg)
lazy = os.environ.get("C4_DENSE") is None
for ep in range(int(os.environ.get("C4_EPOCHS", 2))):
    st = {}
    t = time.perf_counter()
    m.fit_epochs(u, i, x, epochs=1, batch_size=4096, lazy=lazy, stats=st)
    torch.cuda.synchronize()
    print(f"epoch {ep}: wall {1e3 * (time.perf_counter() - t):.1f} ms, device {st['device_ms']:.1f} ms", flush=True)
