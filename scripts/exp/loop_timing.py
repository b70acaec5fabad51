"""Where does a DeviceLoop fit spend its host time?  (experiment; C3 shape)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ctypes as C
from prob_matrix_factorization_b200 import _cabi, synth
from prob_matrix_factorization_b200._engine import DeviceLoop, GammaEngine, row_stride
from prob_matrix_factorization_b200.ratings import DeviceRatings

w, (u, i, x) = synth.workload_ratings("c3")
x = x + np.float32(1.0)
K = w.n_factors
dev = torch.device("cuda", 0)
dr = DeviceRatings(u, i, x, w.n_users, w.n_items, dev, row_bytes=4 * row_stride(K))
hyper = {"user_shape": 5.0 + K * 0.3, "user_rate_prior": 5.0, "item_shape": 5.0 + K * 0.3, "item_rate_prior": 5.0}
eng = GammaEngine(dr, K, 0.3, 0.3, None, None, hyper=hyper)
rng = np.random.default_rng(1)
eng.load_means(rng.random((w.n_users, K), dtype=np.float32) + 0.05, rng.random((w.n_items, K), dtype=np.float32) + 0.05,
               np.full(w.n_users, 1.3, np.float32), np.full(w.n_items, 0.9, np.float32))
for _ in range(3):
    eng.sweep(False)
torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    loop = DeviceLoop(dev, 19)
    t.append(time.perf_counter())
    loop.stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.device(dev), torch.cuda.stream(loop.stream):
        _cabi.call("pmf_loop_begin", loop.stream.cuda_stream, C.byref(loop._h)); t.append(time.perf_counter())
        eng.sweep(False); t.append(time.perf_counter())
        _cabi.call("pmf_loop_decide", loop._h, None, 0, 0.0, 0, 19, loop.iter.data_ptr(), loop.history.data_ptr(), loop.stream.cuda_stream)
        t.append(time.perf_counter())
        _cabi.call("pmf_loop_end", loop._h); t.append(time.perf_counter())
    n, _ = loop.run(); t.append(time.perf_counter())
    loop.free(); t.append(time.perf_counter())
    names = ["ctor", "begin", "capture sweep", "decide", "end+instantiate", "run+sync", "free"]
    print(f"rep {rep}: n={n} " + ", ".join(f"{nm} {1e3 * (b - a):.2f} ms" for nm, a, b in zip(names, t[:-1], t[1:])), flush=True)
    t0 = time.perf_counter()
    for _ in range(19):
        eng.sweep(False)
    torch.cuda.synchronize()
    print(f"   host loop of 19 sweeps: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
