"""Timing of dense top-50 scoring at BASELINE config C4 scale (K=100, 230k items); not part of the product."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prob_matrix_factorization_b200.scoring import top_n  # noqa: E402

B = int(os.environ.get("TOPN_USERS", 8192))
M, K, n = 230_000, 100, 50
rng = np.random.default_rng(0)
Fu = torch.from_numpy(rng.gamma(0.3, 1.0, (B, K)).astype(np.float32)).cuda()
Fi = torch.from_numpy(rng.gamma(0.3, 1.0, (M, K)).astype(np.float32)).cuda()
for tensor in (True, False):
    if not tensor and os.environ.get("TOPN_SKIP_EXACT"):
        continue
    top_n(Fu[:256], Fi, n, tensor_cores=tensor)
    torch.cuda.synchronize()
    t = time.perf_counter()
    idx, sc, st = top_n(Fu, Fi, n, tensor_cores=tensor, batch_rows=2048, return_stats=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"TOPN tensor_cores={tensor}: {B} users x {M} items K={K} top-{n}: {dt * 1e3:.1f} ms -> {B / dt:.0f} user-rows/s, "
          f"{2 * B * M * K / dt / 1e12:.1f} TFLOP/s-equivalent, stats={st}", flush=True)
    if tensor:
        ref = idx
    else:
        print("TOPN paths agree:", bool(np.array_equal(ref, idx)))
