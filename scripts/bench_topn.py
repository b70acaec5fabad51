"""Timing of dense top-50 scoring at BASELINE config C4 scale (K=100, 230k items); not part of the product.

    TOPN_USERS=8192 TOPN_MODES=fused,unfused,exact python scripts/bench_topn.py

Prints one line per mode: wall time of scoring.top_n (device-resident factors in, NumPy results out) and the CUDA-event
time of the pmf_topn library call alone (padded factor tables and workspace resident), and whether the modes returned
identical indices.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prob_matrix_factorization_b200 import _cabi  # noqa: E402
from prob_matrix_factorization_b200.scoring import _as_table, top_n  # noqa: E402


def device_ms(Fu, Fi, n, mode, chunk, reps):
    """CUDA-event time of the library calls that score all rows of Fu (no host work in the timed region)."""
    lib = _cabi.load()
    dev = Fu.device
    Tu, K = _as_table(Fu, dev)
    Ti, _ = _as_table(Fi, dev)
    B, M, ld = Tu.shape[0], Ti.shape[0], Tu.shape[1]
    ws_bytes = lib.pmf_topn_workspace_bytes_ex(chunk, M, K, n, mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    idx = torch.empty((B, n), dtype=torch.int32, device=dev)
    sc = torch.empty((B, n), dtype=torch.float32, device=dev)
    best = None
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s0 in range(0, B, chunk):
            e = min(s0 + chunk, B)
            _cabi.call("pmf_topn", Tu[s0:e].data_ptr(), None, e - s0, Ti.data_ptr(), M, K, ld, n, mode, idx[s0:e].data_ptr(),
                       sc[s0:e].data_ptr(), ws.data_ptr(), ws_bytes, None, _cabi.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best

B = int(os.environ.get("TOPN_USERS", 8192))
M = int(os.environ.get("TOPN_ITEMS", 230_000))
K, n = int(os.environ.get("TOPN_K", 100)), int(os.environ.get("TOPN_N", 50))
MODES = {"fused": True, "unfused": "unfused", "exact": False}
modes = os.environ.get("TOPN_MODES", "fused,unfused,exact").split(",")
reps = int(os.environ.get("TOPN_REPS", 3))
if os.environ.get("TOPN_GROWTH"):
    _cabi.call("pmf_tune", b"topn_growth", int(os.environ["TOPN_GROWTH"]))
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
    dms = device_ms(Fu, Fi, n, {"fused": 1, "unfused": 2, "exact": 0}[name], 8192 if name == "fused" else 2048,
                    reps if name != "exact" else 1)
    print(f"TOPN {name}: {B} users x {M} items K={K} top-{n}: wall {best * 1e3:.2f} ms ({B / best:.0f} user-rows/s); library call "
          f"{dms:.3f} ms on the device ({B / dms * 1e3:.0f} user-rows/s, {2 * B * M * K / dms / 1e9:.1f} TFLOP/s-equivalent), "
          f"stats={st}", flush=True)
    if ref is None:
        ref = idx
    else:
        print(f"TOPN {name} agrees with {modes[0]}:", bool(np.array_equal(ref, idx)), flush=True)
