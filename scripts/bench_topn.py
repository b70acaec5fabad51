"""Timing of dense top-50 scoring at BASELINE config C4 scale (K=100, 230k items); not part of the product.

    TOPN_USERS=8192 TOPN_MODES=fused,unfused,exact python scripts/bench_topn.py

Prints one line per mode (wall time of scoring.top_n from device-resident factors, results copied back) and whether
the modes returned identical indices.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prob_matrix_factorization_b200.scoring import top_n  # noqa: E402

B = int(os.environ.get("TOPN_USERS", 8192))
M = int(os.environ.get("TOPN_ITEMS", 230_000))
K, n = int(os.environ.get("TOPN_K", 100)), int(os.environ.get("TOPN_N", 50))
MODES = {"fused": True, "unfused": "unfused", "exact": False}
modes = os.environ.get("TOPN_MODES", "fused,unfused,exact").split(",")
reps = int(os.environ.get("TOPN_REPS", 3))
rng = np.random.default_rng(0)
Fu = torch.from_numpy(rng.gamma(0.3, 1.0, (B, K)).astype(np.float32)).cuda()
Fi = torch.from_numpy(rng.gamma(0.3, 1.0, (M, K)).astype(np.float32)).cuda()
ref = None
for name in modes:
    tensor = MODES[name]
    kw = {} if name == "fused" else {"batch_rows": 2048}
    top_n(Fu[:256], Fi, n, tensor_cores=tensor, **kw)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps if name != "exact" else 1):
        t = time.perf_counter()
        idx, sc, st = top_n(Fu, Fi, n, tensor_cores=tensor, return_stats=True, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    print(f"TOPN {name}: {B} users x {M} items K={K} top-{n}: {best * 1e3:.2f} ms -> {B / best:.0f} user-rows/s, "
          f"{2 * B * M * K / best / 1e12:.1f} TFLOP/s-equivalent, stats={st}", flush=True)
    if ref is None:
        ref = idx
    else:
        print(f"TOPN {name} agrees with {modes[0]}:", bool(np.array_equal(ref, idx)), flush=True)
