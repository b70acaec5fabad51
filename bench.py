#!/usr/bin/env python
"""Benchmark of the per-iteration training hot path (BASELINE.json metric: rating-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c3|c2] [--impl reference]

A "step" is one full CAVI sweep (user pass + item pass) over every rating of the workload.  The
default workload is BASELINE.json configs[4] ("c5": hpf_cavi K=64, 2M users x 500k items x 100M
ratings) -- the configuration the metric and the 8-GPU target are quoted on; it fits one B200
(2.3 GB of ratings + 1.9 GB of state), so the same workload runs at N = 1, 2, 4, 8 (strong scaling:
ratings sharded by nonzero, factors replicated).  One JSON line is printed by rank 0.

 * value      : nnz * K / device time of K sweeps (CUDA events, barrier+sync both sides, max over ranks),
                inputs resident in HBM.  Working set (>= 2.9 GB) exceeds the 126 MB L2.
 * e2e        : the same metric through the public drop-in API (HPF_CAVI.fit_arrays) from HOST buffers:
                H2D of ratings and initial factors, device grouping (CSR/CSC build), K sweeps and the D2H
                read of E_theta / E_beta all inside the timed region.
 * roofline   : algorithmic bytes of a sweep / measured pass-kernel time, against the
                measured HBM copy bandwidth in MEASURED_PEAKS.json.
 * cpu_baseline / --impl reference : the oracle's C port of the reference algorithm (OpenMP over rows) on
                the box's host cores, bounded sample.  The reference itself is pure Python and absent from
                the GPU box; its row-loop style is timed too (python_rowloop) on a C1-shaped sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

HPF_HP = dict(a=0.3, a_prime=5.0, b_prime=5.0, c=0.3, c_prime=5.0, d_prime=5.0)   # best_hyperparams.txt:5
POISSON_HP = dict(a0=0.1, b0=0.5)                                                  # best_hyperparams.txt:4
METRIC = "rating-updates/sec (nnz*iters/s)"
UNIT = "nnz*iters/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_OUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries also write there (NCCL prints its version banner to fd 1
    when a communicator is created), so fd 1 is pointed at stderr for everything but the result line."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT_OUT


def emit(line):
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_workload(name, sample_nnz=None):
    """The named workload, or a proportionally scaled-down instance of it (users, items and ratings all
    divided by nnz/sample_nnz, so row lengths -- the per-row vs per-rating work split -- are preserved)."""
    import dataclasses
    from prob_matrix_factorization_b200 import synth
    w = synth.WORKLOADS[name]
    if sample_nnz is not None and sample_nnz < w.nnz:
        f = w.nnz / sample_nnz
        w = dataclasses.replace(w, n_users=max(2, int(w.n_users / f)), n_items=max(2, int(w.n_items / f)), nnz=int(sample_nnz))
    nnz = w.nnz
    t = time.time()
    u, i, x = synth.make_ratings(w.n_users, w.n_items, nnz, w.seed)
    if w.model in ("hpf_cavi", "hpf_pytorch"):
        x = x + np.float32(1.0)                       # compare_models.py:180-185 (+1 shift for HPF)
    log(f"[bench] generated {name}: N={w.n_users} M={w.n_items} nnz={nnz} K={w.n_factors} in {time.time() - t:.1f}s")
    return w, u, i, x


def make_model(w, steps, device=None, shard=None, seg_len=None):
    kw = {} if seg_len is None else {"seg_len": seg_len}
    if w.model == "poisson_mf":
        from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
        m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=w.n_factors, max_iter=steps, tol=None, random_state=42,
                                              verbose=False, **POISSON_HP), device=device, shard=shard, **kw)
    else:
        from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
        m = HPF_CAVI(HPF_CAVI_Config(n_factors=w.n_factors, max_iter=steps, tol=None, random_state=42, verbose=False,
                                     **HPF_HP), device=device, shard=shard, **kw)
    m.n_users, m.n_items = w.n_users, w.n_items
    m._auto_close = False            # keep the (peer-mapped) engine alive for the device-resident timing
    return m


def initial_state_f32(m):
    """The reference's own PCG64 draws (bit-identical order), kept as float32 expectations only."""
    t = time.time()
    init = m._initial_state()
    for k in list(init):
        if isinstance(init[k], np.ndarray) and init[k].ndim == 2 and not k.startswith("E_"):
            init[k] = init[k][:1]            # shape/rate draws are not needed after E = shape/rate
    for k in ("E_theta", "E_beta", "E_xi", "E_eta"):
        if k in init:
            init[k] = np.ascontiguousarray(init[k], dtype=np.float32)
    log(f"[bench] host init draws (NumPy PCG64, reference order) in {time.time() - t:.1f}s")
    return init


def pin(arr):
    """Page-lock a NumPy array in place so H2D copies from it are asynchronous DMA."""
    import torch
    if arr.nbytes:
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        if int(rc) != 0:
            log(f"[bench] cudaHostRegister failed rc={rc}; copies fall back to pageable memory")
    return arr


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpus = set(gpu_indices)
        self.path = tempfile.mktemp(prefix="pmf_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception as e:  # nvidia-smi missing
            log(f"[bench] clock sampler unavailable: {e}")
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for ln in f:
                p = [c.strip() for c in ln.split(",")]
                if len(p) < 8 or not p[0].isdigit() or int(p[0]) not in self.gpus:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
                except ValueError:
                    continue
                for n, v in zip(names, p[4:8]):
                    if v == "Active":
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "power_w_max": float(np.max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle; rank 0 only)
# ------------------------------------------------------------------------------------------------
def cpu_port_run(w, sample_nnz, sweeps, threads=0):
    """Oracle C port (OpenMP over rows) on a bounded sample of the workload.  Returns (nnz*it/s, info)."""
    from oracle import c_oracle as CO
    from oracle import pmf_oracle as O
    w, u, i, x = make_workload(w.name, sample_nnz)
    K, N, M = w.n_factors, w.n_users, w.n_items
    threads = threads or CO.max_threads()
    if w.model == "poisson_mf":
        init = O.poisson_init(N, M, K, POISSON_HP["a0"], POISSON_HP["b0"], 42)
        res = CO.poisson_sweeps(u, i, x, N, M, K, POISSON_HP["a0"], POISSON_HP["b0"], sweeps, init["E_theta"],
                                init["E_beta"], threads)
    else:
        init = O.hpf_init(N, M, K, HPF_HP, 42)
        res = CO.hpf_sweeps(u, i, x, N, M, K, HPF_HP, sweeps, init, threads)
    secs = res["sweep_seconds"]
    return len(x) * sweeps / secs, {"cores": threads, "seconds": secs, "sample_nnz": int(len(x)), "sweeps": sweeps,
                                    "shape": f"{N} users x {M} items x {len(x)} ratings"}


def python_rowloop_run():
    """The reference's own style (one NumPy row loop per pass, single core) on a C1-shaped sample."""
    from oracle import pmf_oracle as O
    from prob_matrix_factorization_b200 import synth
    N, M, nnz, K = 20_000, 10_000, 200_000, 50
    u, i, x = synth.make_ratings(N, M, nnz, 20261)
    u64, i64, x64 = u.astype(np.int64), i.astype(np.int64), x.astype(np.float64) + 1.0
    st = O.hpf_init(N, M, K, HPF_HP, 42)
    rp_u, pm_u = O.group_observations(u64, N)
    rp_i, pm_i = O.group_observations(i64, M)
    t = time.time()
    a, b = O.gamma_row_pass(rp_u, pm_u, i64, x64, st["E_theta"], st["E_beta"], HPF_HP["a"], st["E_xi"])
    st["E_theta"] = a / b
    a, b = O.gamma_row_pass(rp_i, pm_i, u64, x64, st["E_beta"], st["E_theta"], HPF_HP["c"], st["E_eta"])
    secs = time.time() - t
    return {"value": nnz / secs, "unit": UNIT, "cores": 1, "sample": f"1 sweep, hpf K={K}, {N}x{M}x{nnz} (C1 shape)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from prob_matrix_factorization_b200 import synth
    w = synth.WORKLOADS[args.workload]
    sample = min(w.nnz, args.cpu_sample)
    steps = max(1, args.steps)
    # warm-up sweeps are folded into one call: the C port has no caches to warm beyond the first touch
    if args.warmup > 0:
        cpu_port_run(w, min(sample, 1_000_000), 1)
    value, info = cpu_port_run(w, sample, steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"] / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{w.name}: {w.model} K={w.n_factors}, {w.n_users} users x {w.n_items} items x "
                                   f"{w.nnz} ratings; timed on a 1/{max(1, round(w.nnz / info['sample_nnz']))} scale "
                                   f"instance ({info['shape']})"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port",
                             "sample": f"{steps} sweep(s) over a scaled-down instance of {w.name} ({info['shape']}, same "
                                       f"row-length distribution), oracle C port (float64), OpenMP over rows"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=["c2", "c3", "c5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="ratings in the CPU-baseline sample")
    ap.add_argument("--seg-len", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default=None, choices=["nccl", "mc"], help="multi-GPU combine of the item pass (default mc)")
    ap.add_argument("--tune", default="", help="comma list key=value passed to pmf_tune (experiments)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from prob_matrix_factorization_b200 import _cabi
    from prob_matrix_factorization_b200.parallel import init_process_group

    if args.exchange:
        os.environ["PMF_EXCHANGE"] = args.exchange
    rank, world, local = init_process_group()
    if world != args.gpus:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _cabi.call("pmf_tune", k.encode(), int(v))
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    shard = (rank, world) if world > 1 else None

    w, u, i, x = make_workload(args.workload)
    pin(u); pin(i); pin(x)
    m = make_model(w, steps, dev, shard, args.seg_len)
    init = initial_state_f32(m)
    for k in ("E_theta", "E_beta", "E_xi", "E_eta"):
        if k in init:
            pin(init[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: public API from host buffers ------------------------------------------------------
    out_t = torch.empty((w.n_users, w.n_factors), dtype=torch.float32, pin_memory=True)
    out_b = torch.empty((w.n_items, w.n_factors), dtype=torch.float32, pin_memory=True)
    e2e = None
    m.config.max_iter = 1
    m.fit_arrays(u, i, x, init)                      # untimed first call: CUDA context, allocator, NCCL
    m._engine.download_means(out_t, out_b, owned_only=world > 1)
    if not args.no_e2e:
        m.config.max_iter = steps
        barrier()
        t0 = time.perf_counter()
        m.fit_arrays(u, i, x, init)
        d2h = m._engine.download_means(out_t, out_b, owned_only=world > 1)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        init_bytes = sum(init[k].nbytes for k in ("E_theta", "E_beta", "E_xi", "E_eta") if k in init)
        e2e = {"value": w.nnz * steps / dt, "unit": UNIT,
               # N > 1: every rank uploads 1/N of the ratings and of the initial factors (the rest travels over
               # NVLink) and reads back the factor rows it owns, so the job moves each byte over PCIe once
               "h2d_bytes_per_step": (12 * w.nnz + init_bytes) / steps,
               "d2h_bytes_per_step": (out_t.numel() + out_b.numel()) * 4 / steps,
               "seconds": dt, "sweeps": steps,
               "includes": "H2D ratings + initial factors (pinned), device CSR+CSC build, sweeps, D2H E_theta/E_beta",
               "excludes": "host NumPy PCG64 draws of the initial state (identical work in the reference)"}
        log(f"[bench] e2e {steps} sweeps from host buffers: {dt * 1e3:.1f} ms")
    eng = m._engine

    # ---- device-resident timing -----------------------------------------------------------------
    sampler = ClockSampler(range(torch.cuda.device_count()) if world > 1 else [local])
    if rank == 0:
        sampler.start()
    # clock-sampling pre-roll: nvidia-smi samples every 50 ms, a short timed region (K sweeps of ~1-8 ms)
    # could end before the first sample, so the same sweeps run untimed for ~0.4 s first; the sampler
    # stays on through warm-up and the timed region (all under the identical load).
    t_pre = time.perf_counter()
    n_pre = 0
    while True:
        eng.sweep(False)
        n_pre += 1
        if n_pre % 4 == 0:
            torch.cuda.synchronize()
            flag = torch.tensor([1.0 if time.perf_counter() - t_pre > 0.4 else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)      # all ranks leave the pre-roll together
            if flag.item() > 0:
                break
    for _ in range(warmup):
        eng.sweep(False)
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for s in range(steps):
        wp = s == steps - 1                       # as in fit(): only the last sweep materialises the Gamma shape/rate tables
        ev[s][0].record()
        eng.user_pass(wp)
        ev[s][1].record()
        eng.item_pass(wp)
        ev[s][2].record()
    stop.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = start.elapsed_time(stop)
    t_user = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t_item = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    if world > 1:
        tt = torch.tensor([ms, t_user, t_item], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, t_user, t_item = (float(v) for v in tt.tolist())
    value = w.nnz * steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (gamma_pass_kernel; both passes of a sweep) -------------
    peak, peak_src = measured_peaks()
    alg_bytes = eng.algorithmic_bytes_per_sweep() / world          # per GPU per sweep
    pass_ms = t_user + t_item                                       # includes the NCCL row exchange when N > 1
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
    # DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the two pass launches of one sweep, from the ncu
    # --set full capture of this command committed as profiles/r1_gamma_pass_final2_ncu_full.csv: user pass
    # 10.4 + 1.6 GB, item pass 20.1 + 0.4 GB.  Below the algorithmic bytes because the 128 MB item table is half
    # L2-resident.  Only known for the configuration that was captured.
    traffic = 32.5e9 if (w.name == "c5" and world == 1) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": "DRAM bytes per sweep (both pass launches), ncu capture in profiles/; "
                "achieved counts algorithmic bytes per sweep the same way",
                "peak_source": peak_src, "kernel": "pmf::gamma_pass_kernel (+gamma_multi_kernel)",
                "algorithmic_bytes_per_sweep_per_gpu": alg_bytes, "user_pass_ms": t_user, "item_pass_ms": t_item,
                "bytes_per_rating_update": eng.algorithmic_bytes_per_sweep() / w.nnz}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline:
        try:
            sample = min(w.nnz, args.cpu_sample)
            v, info = cpu_port_run(w, sample, 1)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port",
                   "sample": f"1 sweep over a scaled-down instance of {w.name} ({info['shape']}, same row-length "
                             f"distribution) by the oracle's C port (float64, OpenMP over rows), {info['seconds']:.1f}s",
                   "python_rowloop": python_rowloop_run()}
        except Exception as e:  # the baseline is reporting only; never lose the GPU line to it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    ws_gb = (2 * w.nnz * 8 + (w.n_users + w.n_items) * eng.ld * 4 * 3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{w.name}: {w.model} K={w.n_factors}, {w.n_users} users x {w.n_items} items x "
                                   f"{w.nnz} ratings (BASELINE.json configs[{int(w.name[1]) - 1}])",
                       "sharding": "ratings by nonzero along nnz-balanced user ranges; E_theta rows live with their owner, E_beta replicated" if world > 1 else "single GPU",
                       "combine": {"mc": "item pass: per-rank row sums added in the NVSwitch (multimem.ld_reduce) by the row's "
                                         "owner, Gamma update, new rows replicated with multimem.st; item rows in "
                                         f"{eng.item_chunks} chunks, combine of a chunk overlaps the pass over the next",
                                   "nccl": "item pass: NCCL all-reduce of the row sums, every rank updates every row",
                                   "none": None}[eng.exchange],
                       "tiles": {"user_pass": len(eng.r.user_tiles), "item_pass": len(eng.r.item_tiles)},
                       "l2": (f"working set {ws_gb:.2f} GB (ratings + factor tables) vs 126 MB L2: "
                              + ("inputs larger than L2, no flush" if ws_gb > 0.5 else
                                 "comparable to L2 -- tables stay L2-resident between sweeps, as they do in a real fit; "
                                 "no flush (a flush would time a cold start no training loop sees)")),
                       "seg_len": eng.r.seg_len},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": eng.launches_per_sweep * steps, "clocks": clocks}
    if clocks is not None:
        clocks["window"] = "pre-roll + warm-up + timed region (same sweeps), nvidia-smi every 50 ms"
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
