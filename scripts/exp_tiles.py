"""Tile-count sweep of the two passes at BASELINE config C5 (experiment; numbers go to profiles/README.md).

    python scripts/exp_tiles.py "1x1,1x8,2x8,2x16,3x8" [workload]

Each combo UxI = user-pass tiles (item ranges) x item-pass tiles (user ranges).  The rating list is generated and
uploaded once; per combo the tiles are rebuilt and 10 sweeps are timed with CUDA events (after 3 warm-up sweeps)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from prob_matrix_factorization_b200 import _cabi, synth  # noqa: E402
from prob_matrix_factorization_b200._engine import GammaEngine, row_stride  # noqa: E402
from prob_matrix_factorization_b200.ratings import DeviceRatings  # noqa: E402

combos = (sys.argv[1] if len(sys.argv) > 1 else "1x1,1x8,2x8").split(",")
name = sys.argv[2] if len(sys.argv) > 2 else "c5"
w, (u, i, x) = synth.workload_ratings(name)
x = x + np.float32(1.0)
K = w.n_factors
dev = torch.device("cuda", 0)
ud, idv, xd = (torch.from_numpy(a).to(dev) for a in (u, i, x))
rng = np.random.default_rng(1)
Et = rng.random((w.n_users, K), dtype=np.float32) + 0.05
Eb = rng.random((w.n_items, K), dtype=np.float32) + 0.05
hyper = {"user_shape": 5.0 + K * 0.3, "user_rate_prior": 5.0, "item_shape": 5.0 + K * 0.3, "item_rate_prior": 5.0}
for combo in combos:
    combo, _, tune = combo.partition(":")
    for kv in filter(None, tune.split("+")):
        k, v = kv.split("=")
        _cabi.call("pmf_tune", k.encode(), int(v))
    tu, ti = (int(v) for v in combo.split("x"))
    dr = DeviceRatings(ud, idv, xd, w.n_users, w.n_items, dev, row_bytes=4 * row_stride(K), user_pass_tiles=tu, item_pass_tiles=ti)
    eng = GammaEngine(dr, K, 0.3, 0.3, None, None, hyper=hyper)
    eng.load_means(Et, Eb, np.full(w.n_users, 1.3, np.float32), np.full(w.n_items, 0.9, np.float32))
    for _ in range(3):
        eng.sweep(False)
    R = 10
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(R)]
    torch.cuda.synchronize()
    for s in range(R):
        ev[s][0].record(); eng.user_pass(False); ev[s][1].record(); eng.item_pass(False); ev[s][2].record()
    torch.cuda.synchronize()
    tu_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev])); ti_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    print(f"TILES {name} {combo}{':' + tune if tune else ''}: user pass {tu_ms:.3f} ms, item pass {ti_ms:.3f} ms, sweep {tu_ms + ti_ms:.3f} ms "
          f"-> {w.nnz / (tu_ms + ti_ms) * 1e3:.3e} nnz*it/s (seg_len {dr.seg_len_user}/{dr.seg_len_item}, launches {eng.launches_per_sweep})", flush=True)
    for kv in filter(None, tune.split("+")):
        _cabi.call("pmf_tune", kv.split("=")[0].encode(), -1 if "chunk_reduce" in kv else 0)
    dr.free()
    del eng, dr
    torch.cuda.empty_cache()
