"""Host-side planning logic that needs no GPU: tile bounds, per-launch segment length, bench.py's helpers."""
import importlib.util
import json
import os

import numpy as np

from conftest import REPO
from prob_matrix_factorization_b200.ratings import auto_seg_len, tile_bounds


def test_tile_bounds_cover_the_range_once():
    mb = 1 << 20
    for lo, hi, row_bytes, tile_mb, want in [(0, 2_000_000, 256, 128, 4), (0, 500_000, 256, 128, 1), (0, 750_000, 256, 128, 1),
                                             (0, 800_000, 256, 128, 2), (250_000, 500_000, 256, 128, 1), (0, 10, 256, 128, 1),
                                             (0, 0, 256, 128, 1), (0, 1000, None, 128, 1)]:
        b = tile_bounds(lo, hi, row_bytes, tile_mb * mb)
        assert len(b) - 1 == want, (lo, hi, b)
        assert b[0] == lo and b[-1] == hi and np.all(np.diff(b) >= 0)
    b = tile_bounds(100, 1100, 256, n_tiles=7)                      # explicit tile count
    assert len(b) == 8 and b[0] == 100 and b[-1] == 1100 and np.diff(b).max() - np.diff(b).min() <= 1
    assert len(tile_bounds(0, 3, 256, n_tiles=10)) - 1 == 3         # never more tiles than rows


def test_auto_seg_len_is_per_launch_and_clamped():
    assert auto_seg_len(100) == 64 and auto_seg_len(1_100_000) == 64
    assert auto_seg_len(100_000_000) == 1024 and auto_seg_len(10 ** 10) == 1024
    assert auto_seg_len(12_500_000) == 384                          # one rank's item pass at 8 GPUs
    assert all(auto_seg_len(n) % 8 == 0 for n in (1, 5_000_000, 33_000_000))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_helpers(monkeypatch):
    b = _bench()
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    assert b.rank_gpu_indices(4) == [0, 1, 2, 3]
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6,7")
    assert b.rank_gpu_indices(2) == [4, 5]
    monkeypatch.setenv("OMP_NUM_THREADS", "1")                       # what torchrun exports: must not shrink the CPU legs
    assert b.host_cores() == len(os.sched_getaffinity(0)) >= 1
    w = b.workload_spec("c3+elbo")
    assert (w.model, w.n_factors, w.nnz) == ("hpf_cavi", 50, 1_100_000) and "ELBO" in b.describe(w, "c3+elbo")
    traffic, src = b.ncu_traffic("c5/n1/tiles1x4")
    with open(os.path.join(REPO, "profiles", "dram_traffic.json")) as f:
        assert traffic == json.load(f)["c5/n1/tiles1x4"]["bytes_per_step"] and "ncu" in src
    assert b.ncu_traffic("no/such/config") == (None, None)
