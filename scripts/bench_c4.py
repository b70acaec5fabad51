"""BASELINE config C4 timing (HPF_PyTorch K=100, 200k x 230k x 1.1M ratings): epochs through the drop-in.
Not part of the product; numbers go to profiles/README.md."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prob_matrix_factorization_b200 import synth  # noqa: E402
from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config  # noqa: E402

w, (u, i, x) = synth.workload_ratings("c4")
x = x + 1.0
N, M, K, nnz = w.n_users, w.n_items, w.n_factors, w.nnz
uc = np.bincount(u, minlength=N); ic = np.bincount(i, minlength=M)
cfg = HPF_PyTorch_Config(n_factors=K, a=0.3, c=0.3, lr=5e-4)
for mode in ("lazy", "dense"):
    torch.manual_seed(0)
    m = HPF_PyTorch(N, M, uc, ic, cfg)
    m.fit_epochs(u, i, x, epochs=1, batch_size=4096, lazy=mode == "lazy")
    torch.cuda.synchronize(); t = time.perf_counter()
    E = 5
    losses = m.fit_epochs(u, i, x, epochs=E, batch_size=4096, lazy=mode == "lazy")
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"C4 fit_epochs[{mode}]: {E} epochs in {dt * 1e3:.0f} ms -> {nnz * E / dt:.3e} ratings*epochs/s, last loss {losses[-1]:.1f}", flush=True)
# the scripts' own loop (compare_models.py:291-313) on top of the drop-in: DataLoader + torch.optim.Adam
torch.manual_seed(0)
m = HPF_PyTorch(N, M, uc, ic, cfg)
opt = torch.optim.Adam(m.parameters(), lr=cfg.lr)


class DS(torch.utils.data.Dataset):
    def __init__(s): s.u = torch.LongTensor(u.astype(np.int64)); s.i = torch.LongTensor(i.astype(np.int64)); s.r = torch.FloatTensor(x)
    def __len__(s): return len(s.r)
    def __getitem__(s, k): return s.u[k], s.i[k], s.r[k]


loader = torch.utils.data.DataLoader(DS(), batch_size=4096, shuffle=True)
t = time.perf_counter(); tot = 0.0
for users, items, ratings in loader:
    opt.zero_grad(); loss = m.loss(users, items, ratings); loss.backward(); opt.step(); tot += loss.item()
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"C4 script loop (DataLoader + torch Adam on the drop-in): 1 epoch in {dt:.2f} s -> {nnz / dt:.3e} ratings*epochs/s", flush=True)
