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
for rep, E in enumerate((1, 1, int(os.environ.get("C4_EPOCHS", 8)))):
    st = {}
    t = time.perf_counter()
    m.fit_epochs(u, i, x, epochs=E, batch_size=4096, lazy=lazy, stats=st)
    torch.cuda.synchronize()
    print(f"call {rep}: {E} epoch(s): wall {1e3 * (time.perf_counter() - t) / E:.1f} ms/epoch, device {st['device_ms'] / E:.2f} ms/epoch", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.fit_epochs(u, i, x, epochs=4, batch_size=4096, lazy=lazy, shuffle=False)
e1.record(); torch.cuda.synchronize()
print(f"4 unshuffled epochs (no host permutation): {e0.elapsed_time(e1) / 4:.2f} ms/epoch on the device", flush=True)
