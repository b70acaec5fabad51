#!/usr/bin/env python
"""Benchmark of the per-iteration training hot path (BASELINE.json metric: rating-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c3|c3+elbo|c2|c1|c4|topn] [--impl reference]

Default workload: BASELINE.json configs[4] ("c5": hpf_cavi K=64, 2M users x 500k items x 100M ratings) -- the
configuration the metric and the 8-GPU target are quoted on; it fits one B200, so the same workload runs at
N = 1, 2, 4, 8 (strong scaling: ratings sharded by nonzero along user ranges).  The other workloads are the remaining
BASELINE configs (c1 gaussian_mf, c2 poisson_mf, c3 hpf_cavi [+ELBO every sweep], c4 HPF-MAP epochs, topn = the
dense top-50 scoring of configs[3]); c1 / c4 / topn run on one GPU (N > 1: rank 0 alone works and reports).
One JSON line is printed by rank 0.

CAVI workloads (c2, c3, c3+elbo, c5) -- a "step" is one full sweep (user pass + item pass) over every rating:
 * value      : nnz * K / device time of K sweeps (CUDA events, barrier+sync both sides, max over ranks), inputs
                resident in HBM; as in fit(), the last sweep of the K also writes the Gamma shape/rate tables.
 * e2e        : the same metric through the public drop-in API (fit_arrays) from pinned HOST buffers: H2D of ratings
                and initial factors, routing + device grouping (CSR/CSC build), K sweeps, D2H of E_theta / E_beta.
 * e2e_fit_df : the reference's own call, Model(config).fit(DataFrame), including the host-side NumPy PCG64 draws of
                the initial state (bit-identical to the reference's) and the DataFrame -> array conversions.
 * roofline   : algorithmic bytes of a sweep / measured pass time, against MEASURED_PEAKS.json.
 * parity_check : N > 1: 2 sweeps of the sharded engine against 2 sweeps of a single-GPU engine on rank 0 from the same
                initial state (max-norm relative error) + replica checksums equal on all ranks;  N = 1: 1 sweep of
                the engine on the cpu_baseline sample against the oracle's C port.
 * cpu_baseline / --impl reference : the oracle's C port of the reference algorithm (OpenMP over rows) on ALL of the
                box's host cores (the process's CPU affinity, never the OMP_NUM_THREADS=1 torchrun exports), bounded sample.
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
GAUSS_HP = dict(sigma2=0.5, eta_theta2=0.1, eta_beta2=0.1, eta_bias2=0.1)          # best_hyperparams.txt:3
MAP_HP = dict(a=0.3, c=0.3, lr=5e-4)                                               # best_hyperparams.txt:6
METRIC = "rating-updates/sec (nnz*iters/s)"
UNIT = "nnz*iters/s"
WORKLOAD_ALIASES = {"c3+elbo": "c3"}


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


def host_cores():
    """Cores this process may run on -- what the CPU legs use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1500.0, "bf16_tflops_sustained": 1500.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """DRAM bytes per step of the dominant kernel(s) from the committed ncu --set full capture of this command
    (profiles/dram_traffic.json, one entry per captured configuration) or None."""
    p = os.path.join(REPO, "profiles", "dram_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    e = d.get(key)
    return (e["bytes_per_step"], e.get("source")) if e else (None, None)


def workload_spec(name):
    from prob_matrix_factorization_b200 import synth
    return synth.WORKLOADS[WORKLOAD_ALIASES.get(name, name)]


def make_workload(name, sample_nnz=None):
    """The named workload, or a proportionally scaled-down instance of it (users, items and ratings all
    divided by nnz/sample_nnz, so row lengths -- the per-row vs per-rating work split -- are preserved)."""
    import dataclasses
    from prob_matrix_factorization_b200 import synth
    w = workload_spec(name)
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


def describe(w, name):
    idx = int(w.name[1]) - 1
    extra = " + ELBO after every sweep" if name == "c3+elbo" else ""
    return (f"{name}: {w.model} K={w.n_factors}, {w.n_users} users x {w.n_items} items x {w.nnz} ratings{extra} "
            f"(BASELINE.json configs[{idx}])")


def make_model(w, steps, device=None, shard=None, seg_len=None, track_elbo=False):
    kw = {} if seg_len is None else {"seg_len": seg_len}
    if w.model == "poisson_mf":
        from prob_matrix_factorization_b200.poisson_mf_cavi import PoissonMFCAVI, PoissonMFCAVIConfig
        m = PoissonMFCAVI(PoissonMFCAVIConfig(n_factors=w.n_factors, max_iter=steps, tol=None, random_state=42,
                                              verbose=False, **POISSON_HP), device=device, shard=shard, **kw)
    else:
        from prob_matrix_factorization_b200.hpf_cavi import HPF_CAVI, HPF_CAVI_Config
        m = HPF_CAVI(HPF_CAVI_Config(n_factors=w.n_factors, max_iter=steps, tol=None, random_state=42, verbose=False,
                                     **HPF_HP), device=device, shard=shard, track_elbo=track_elbo, **kw)
    m.n_users, m.n_items = w.n_users, w.n_items
    m._auto_close = False            # keep the (symmetric-memory) engine alive for the device-resident timing
    return m


def initial_state_f32(m, keep_params=False):
    """The reference's own PCG64 draws (bit-identical order); expectations as float32."""
    t = time.time()
    init = m._initial_state()
    if not keep_params:
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


def rank_gpu_indices(world):
    """nvidia-smi indices of the GPUs the ranks of this job run on (local rank r = r-th visible device)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if all(v.isdigit() for v in ids):
            return [int(v) for v in ids[:world]]
    return list(range(world))


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
                "samples": len(sm), "gpus": sorted(self.gpus), "reasons": sorted(reasons),
                "window": "pre-roll + warm-up + timed region (same work), nvidia-smi every 50 ms, the ranks' GPUs only"}


class Job:
    """Process-group / device context shared by the GPU workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from prob_matrix_factorization_b200 import _cabi
        from prob_matrix_factorization_b200.parallel import init_process_group
        self.torch, self.dist = torch, dist
        if args.exchange:
            os.environ["PMF_EXCHANGE"] = args.exchange
        self.rank, self.world, self.local = init_process_group()
        if self.world != args.gpus:
            log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={self.world}; using WORLD_SIZE")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        for kv in filter(None, args.tune.split(",")):
            k, v = kv.split("=")
            _cabi.call("pmf_tune", k.encode(), int(v))
        self.steps, self.warmup = max(1, args.steps), max(3, args.warmup)
        self.args = args

    def my_gpus(self):
        """nvidia-smi indices to sample: every rank's GPU when N > 1, this process's GPU otherwise."""
        if self.world > 1:
            return rank_gpu_indices(self.world)
        vis = rank_gpu_indices(max(self.torch.cuda.device_count(), 1))
        return vis[self.local:self.local + 1]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def preroll(self, fn, seconds=0.4):
        """nvidia-smi samples every 50 ms; a short timed region could end before the first sample, so the same work runs
        untimed for ~0.4 s first (the sampler stays on through warm-up and the timed region)."""
        torch = self.torch
        t0, n = time.perf_counter(), 0
        while True:
            fn()
            n += 1
            if n % 4 == 0:
                torch.cuda.synchronize()
                flag = torch.tensor([1.0 if time.perf_counter() - t0 > seconds else 0.0], device=self.dev)
                if self.world > 1:
                    self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)      # all ranks leave the pre-roll together
                if flag.item() > 0:
                    break

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle; rank 0 only)
# ------------------------------------------------------------------------------------------------
def cpu_port_run(name, sample_nnz, sweeps, threads=0, want_state=False):
    """Oracle C port (OpenMP over rows) on a bounded sample of the workload.  Returns (nnz*it/s, info)."""
    from oracle import c_oracle as CO
    from oracle import pmf_oracle as O
    w, u, i, x = make_workload(name, sample_nnz)
    K, N, M = w.n_factors, w.n_users, w.n_items
    threads = threads or host_cores()
    if w.model == "poisson_mf":
        init = O.poisson_init(N, M, K, POISSON_HP["a0"], POISSON_HP["b0"], 42)
        res = CO.poisson_sweeps(u, i, x, N, M, K, POISSON_HP["a0"], POISSON_HP["b0"], sweeps, init["E_theta"],
                                init["E_beta"], threads)
    elif w.model == "gaussian_mf":
        init = O.gauss_init(N, M, K, 42)
        res = CO.gauss_sweeps(u, i, x.astype(np.float64) - float(x.mean()), N, M, K, GAUSS_HP["sigma2"],
                              GAUSS_HP["eta_theta2"], GAUSS_HP["eta_beta2"], GAUSS_HP["eta_bias2"], sweeps, init, True, threads)
    else:
        init = O.hpf_init(N, M, K, HPF_HP, 42)
        res = CO.hpf_sweeps(u, i, x, N, M, K, HPF_HP, sweeps, init, threads)
    secs = res["sweep_seconds"]
    elbo_secs = 0.0
    if name == "c3+elbo":          # the reference has no ELBO; the oracle's NumPy restatement, once per sweep
        t = time.time()
        O.hpf_elbo(u, i, x, dict(res, gamma_a_xi=init["gamma_a_xi"], gamma_a_eta=init["gamma_a_eta"]), HPF_HP)
        elbo_secs = (time.time() - t) * sweeps
        secs += elbo_secs
    info = {"cores": threads, "seconds": secs, "sample_nnz": int(len(x)), "sweeps": sweeps, "elbo_seconds": elbo_secs,
            "shape": f"{N} users x {M} items x {len(x)} ratings"}
    if want_state:
        info["state"] = (w, u, i, x, init, res)
    return len(x) * sweeps / secs, info


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


def cpu_map_run(steps):
    """HPF-MAP (c4) on the host: the oracle's NumPy restatement of loss+backward (hpf_pytorch.py:71-184) and of torch's
    dense Adam, `steps` mini-batches of 4096 on the full-size tables (float32, as the reference computes)."""
    from oracle import pmf_oracle as O
    w, u, i, x = make_workload("c4")
    N, M, K, B = w.n_users, w.n_items, w.n_factors, 4096
    rng = np.random.default_rng(0)
    P = {"theta": rng.standard_normal((N, K)).astype(np.float32), "beta": rng.standard_normal((M, K)).astype(np.float32),
         "xi": rng.standard_normal(N).astype(np.float32), "eta": rng.standard_normal(M).astype(np.float32)}
    Mo = {k: np.zeros_like(v) for k, v in P.items()}
    Vo = {k: np.zeros_like(v) for k, v in P.items()}
    us = (1.0 / (np.bincount(u, minlength=N).astype(np.float32) + np.float32(1e-6))).astype(np.float32)
    its = (1.0 / (np.bincount(i, minlength=M).astype(np.float32) + np.float32(1e-6))).astype(np.float32)
    cfg = dict(a=MAP_HP["a"], a_prime=0.3, b_prime=1.0, c=MAP_HP["c"], c_prime=0.3, d_prime=1.0)
    perm = rng.permutation(w.nnz)
    t = time.time()
    for s in range(steps):
        b = perm[s * B:(s + 1) * B]
        _, G = O.hpf_map_loss_grads(P, u[b].astype(np.int64), i[b].astype(np.int64), x[b], us, its, cfg, dtype=np.float32)
        O.adam_dense_step(P, G, Mo, Vo, s + 1, MAP_HP["lr"])
    secs = time.time() - t
    return steps * B / secs, {"cores": 1, "seconds": secs, "sample": f"{steps} mini-batches of {B} ratings (dense Adam over all "
                              f"{(N + M) * (K + 1)} parameters per step, NumPy float32)"}


def cpu_topn_run(rows):
    """Dense scoring on the host: float32 U V^T for `rows` users + per-row top-50 (NumPy matmul + argpartition)."""
    M, K, n = 230_000, 100, 50
    rng = np.random.default_rng(0)
    Fu = rng.gamma(0.3, 1.0, (rows, K)).astype(np.float32)
    Fi = rng.gamma(0.3, 1.0, (M, K)).astype(np.float32)
    t = time.time()
    S = Fu @ Fi.T
    part = np.argpartition(-S, n, axis=1)[:, :n]
    sc = np.take_along_axis(S, part, axis=1)
    order = np.argsort(-sc, axis=1, kind="stable")
    np.take_along_axis(part, order, axis=1)
    secs = time.time() - t
    return rows / secs, {"cores": host_cores(), "seconds": secs, "sample": f"{rows} user rows x {M} items, K={K}, top-{n} "
                         "(NumPy float32 GEMM + argpartition)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    name = args.workload
    steps = max(1, args.steps)
    cores = host_cores()
    if name == "c4":
        n = min(steps, 12)
        value, info = cpu_map_run(n)
        metric, unit, sample = "rating-updates/sec (ratings*epochs/s)", "ratings*epochs/s", info["sample"]
        w = workload_spec("c4")
        ms_per_step = 1e3 * info["seconds"] / n * (-(-w.nnz // 4096))      # per epoch, extrapolated from the sample
        cores = info["cores"]
    elif name == "topn":
        value, info = cpu_topn_run(512)
        metric, unit, sample = "user-rows/sec (dense U V^T top-50 over 230k items, K=100)", "user-rows/s", info["sample"]
        ms_per_step = 1e3 * info["seconds"] * 8192 / 512
    else:
        w = workload_spec(name)
        sample_nnz = min(w.nnz, args.cpu_sample)
        if args.warmup > 0:       # the C port has no caches to warm beyond the first touch
            cpu_port_run(name, min(sample_nnz, 1_000_000), 1, cores)
        value, info = cpu_port_run(name, sample_nnz, steps, cores)
        metric, unit = METRIC, UNIT
        scale = max(1, round(w.nnz / info["sample_nnz"]))
        sample = (f"{steps} sweep(s) over a 1/{scale}-scale instance of {name} ({info['shape']}, same row-length "
                  f"distribution), oracle C port (float64), OpenMP over rows on {cores} threads")
        ms_per_step = 1e3 * info["seconds"] / steps
    w = None if name == "topn" else workload_spec(name)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64" if name not in ("c4", "topn") else "f32", "data": "synthetic",
            "config": {"workload": describe(w, name) if w else "topn: 8192 users x 230000 items, K=100, top-50", "sample": sample},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm: CAVI workloads (c2, c3, c3+elbo, c5)
# ------------------------------------------------------------------------------------------------
def parity_single_gpu(job, name, info):
    """N = 1: the engine on the cpu_baseline sample (1 sweep) against the oracle's C port result of the same sweep."""
    w, u, i, x, init, res = info["state"]
    m = make_model(w, 1, job.dev)
    m._auto_close = True
    m.fit_arrays(u, i, x, dict(init))
    worst = 0.0
    for k in ("E_theta", "E_beta"):
        ref = res[k]
        worst = max(worst, float(np.max(np.abs(getattr(m, k) - ref)) / np.max(np.abs(ref))))
    return {"result": "ok" if worst < 1e-5 else "FAIL", "kind": "engine vs oracle C port, 1 sweep on the cpu_baseline sample "
            f"({info['shape']})", "rel_max": worst, "tolerance": 1e-5}


def parity_multi_gpu(job, w, u, i, x, init, m):
    """N > 1: 2 sweeps of the sharded engine vs 2 sweeps of a single-GPU engine (rank 0) from the same initial state;
    replica checksums must agree on all ranks."""
    torch, dist = job.torch, job.dist
    eng = m._engine
    eng.load_means(init["E_theta"], init["E_beta"], init.get("E_xi"), init.get("E_eta"))
    eng.sweep(False)
    eng.sweep(True)
    eng.sync_params()
    torch.cuda.synchronize()
    sums = torch.stack([eng.E_theta.double().sum(), eng.E_beta.double().sum()])
    allsums = [torch.zeros_like(sums) for _ in range(job.world)]
    dist.all_gather(allsums, sums)
    identical = all(torch.equal(allsums[0], t) for t in allsums)
    rel = torch.zeros(2, dtype=torch.float64, device=job.dev)
    if job.rank == 0:
        one = make_model(w, 2, job.dev, None, job.args.seg_len)
        one.fit_arrays(u, i, x, init)
        e1 = one._engine
        for k, (a, b) in enumerate(((eng.E_theta, e1.E_theta), (eng.E_beta, e1.E_beta))):
            rel[k] = (a.double() - b.double()).abs().max() / b.double().abs().max()
        del one, e1
    dist.broadcast(rel, src=0)
    worst = float(rel.max().item())
    ok = identical and worst < 1e-5
    return {"result": "ok" if ok else "FAIL", "kind": f"2 sweeps sharded over {job.world} GPUs vs 2 sweeps on one GPU, same "
            "initial state; E_theta / E_beta max-norm relative error; replica checksums all-gathered",
            "rel_max": worst, "tolerance": 1e-5, "replicas_identical": bool(identical)}


def run_cavi(job, name):
    torch = job.torch
    args, rank, world, dev, steps, warmup = job.args, job.rank, job.world, job.dev, job.steps, job.warmup
    with_elbo = name == "c3+elbo"
    shard = (rank, world) if world > 1 else None
    w, u, i, x = make_workload(name)
    pin(u); pin(i); pin(x)
    m = make_model(w, steps, dev, shard, args.seg_len, track_elbo=with_elbo)
    init = initial_state_f32(m, keep_params=with_elbo)
    for k in ("E_theta", "E_beta", "E_xi", "E_eta"):
        if k in init:
            pin(init[k])

    # ---- e2e: public API from host buffers ------------------------------------------------------
    out_t = torch.empty((w.n_users, w.n_factors), dtype=torch.float32, pin_memory=True)
    out_b = torch.empty((w.n_items, w.n_factors), dtype=torch.float32, pin_memory=True)
    e2e = e2e_df = None
    m.config.max_iter = 2
    m.fit_arrays(u, i, x, init)                      # untimed first call: CUDA context, allocator, streams, NCCL
    m._engine.download_means(out_t, out_b, owned_only=world > 1)
    if not args.no_e2e:
        m.config.max_iter = steps
        job.barrier()
        t0 = time.perf_counter()
        m.fit_arrays(u, i, x, init)
        m._engine.download_means(out_t, out_b, owned_only=world > 1)
        job.barrier()
        dt, = job.max_over_ranks(time.perf_counter() - t0)
        init_bytes = sum(init[k].nbytes for k in ("E_theta", "E_beta", "E_xi", "E_eta") if k in init)
        e2e = {"value": w.nnz * steps / dt, "unit": UNIT,
               # N > 1: every rank uploads 1/N of the ratings, its own rows of E_theta and 1/N of E_beta (the rest travels
               # over NVLink) and reads back the factor rows it owns, so the job moves each byte over PCIe once
               "h2d_bytes_per_step": (12 * w.nnz + init_bytes) / steps,
               "d2h_bytes_per_step": (out_t.numel() + out_b.numel()) * 4 / steps,
               "seconds": dt, "sweeps": steps,
               "includes": "H2D ratings + initial factors (pinned), routing + device CSR/CSC build, sweeps"
                           + (" + ELBO per sweep" if with_elbo else "") + ", D2H E_theta/E_beta",
               "excludes": "host NumPy PCG64 draws of the initial state (see e2e_fit_df)"}
        log(f"[bench] e2e {steps} sweeps from host buffers: {dt * 1e3:.1f} ms")
        if not args.no_fit_df:
            import pandas as pd
            df = pd.DataFrame({"u": u.astype(np.int64), "i": i.astype(np.int64), "rating": x.astype(np.float64)})
            m2 = make_model(w, steps, dev, shard, args.seg_len, track_elbo=with_elbo)
            m2.n_users = m2.n_items = None
            m._engine.close()
            job.barrier()
            t0 = time.perf_counter()
            m2.fit(df)
            m2._engine.download_means(out_t, out_b, owned_only=world > 1)
            job.barrier()
            dt2, = job.max_over_ranks(time.perf_counter() - t0)
            e2e_df = {"value": w.nnz * steps / dt2, "unit": UNIT, "seconds": dt2, "sweeps": steps,
                      "call": "Model(config).fit(DataFrame[u,i,rating] int64/int64/float64) -> E_theta/E_beta on the host",
                      "includes": "everything in e2e + DataFrame column conversions + the reference's PCG64 gamma draws of the "
                                  "initial state on the host (bit-identical order, single NumPy stream)"}
            log(f"[bench] e2e fit(DataFrame) {steps} sweeps: {dt2 * 1e3:.1f} ms")
            m2._engine.close()
            del m2, df
            m.config.max_iter = 1
            m.fit_arrays(u, i, x, init)              # the engine the device-resident timing uses
    eng = m._engine

    # ---- device-resident timing -----------------------------------------------------------------
    def one_step(wp):
        eng.sweep(wp)
        if with_elbo:
            eng.elbo(m.config)

    sampler = ClockSampler(job.my_gpus())
    if rank == 0:
        sampler.start()
    job.preroll(lambda: one_step(with_elbo))
    for _ in range(warmup):
        one_step(with_elbo)
    job.barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.barrier()
    start.record()
    for s in range(steps):
        wp = with_elbo or s == steps - 1           # as in fit(): only the last sweep materialises the Gamma shape/rate tables
        ev[s][0].record()
        eng.user_pass(wp)
        ev[s][1].record()
        eng.item_pass(wp)
        ev[s][2].record()
        if with_elbo:
            eng.elbo(m.config)
        ev[s][3].record()
    stop.record()
    job.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = start.elapsed_time(stop)
    t_user = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t_item = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    t_elbo = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    ms, t_user, t_item, t_elbo = job.max_over_ranks(ms, t_user, t_item, t_elbo)
    pipelined = None
    if world > 1 and not with_elbo and eng.exchange == "mc":
        # what fit() runs on several GPUs: K sweeps software-pipelined (GammaEngine.sweeps: the cross-rank combines of an
        # item pass run under the next sweep's user pass); the pass-by-pass loop above keeps the per-pass split
        eng.sweeps(warmup)
        job.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        eng.sweeps(steps)
        s1.record()
        job.barrier()
        ms_pipe, = job.max_over_ranks(s0.elapsed_time(s1))
        pipelined = {"ms_per_step": ms_pipe / steps, "unpipelined_ms_per_step": ms / steps}
        ms = ms_pipe
    value = w.nnz * steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (gamma_pass_kernel; both passes of a sweep) -------------
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    alg_bytes = eng.algorithmic_bytes_per_sweep() / world          # per GPU per sweep
    pass_ms = ms / steps if pipelined else t_user + t_item           # includes the cross-rank combine when N > 1
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(f"{name}/n{world}/tiles{len(eng.r.user_tiles)}x{len(eng.r.item_tiles)}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": (f"dram__bytes_read+write of the pass launches of one sweep, {traffic_src}"
                                                     if traffic else "no ncu capture of this exact configuration committed; see profiles/README.md"),
                "peak_source": peak_src + " hbm_gbs", "kernel": "pmf::gamma_pass_kernel (+gamma_multi_kernel"
                + (", gamma_combine_kernel)" if world > 1 else ")"),
                "algorithmic_bytes_per_sweep_per_gpu": alg_bytes, "user_pass_ms": t_user, "item_pass_ms": t_item,
                "bytes_per_rating_update": eng.algorithmic_bytes_per_sweep() / w.nnz,
                "note": "achieved counts every gathered factor row as DRAM traffic (SURVEY.md §8d); rows served by the L2 make "
                        "it exceed the DRAM peak -- the L2 itself (~6300 B/clk) then bounds the pass, see DESIGN.md §3.1"}
    if with_elbo:
        roofline["elbo_ms"] = t_elbo
    if pipelined:
        roofline["pipelined"] = pipelined
        roofline["note_passes"] = ("user_pass_ms / item_pass_ms come from a pass-by-pass loop (each item pass waits for its "
                                   "combines); the timed region runs the sweeps pipelined, as fit() does")

    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_multi_gpu(job, w, u, i, x, init, m)

    if rank != 0:
        job.finish()
        return 0

    cpu = None
    if not args.no_cpu_baseline:
        try:
            sample = min(w.nnz, args.cpu_sample)
            v, info = cpu_port_run(name, sample, 1, want_state=world == 1 and not args.no_parity)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port",
                   "sample": f"1 sweep over a 1/{max(1, round(w.nnz / info['sample_nnz']))}-scale instance of {name} "
                             f"({info['shape']}, same row-length distribution) by the oracle's C port (float64, OpenMP "
                             f"over rows), {info['seconds']:.1f}s",
                   "python_rowloop": python_rowloop_run()}
            if "state" in info:
                parity = parity_single_gpu(job, name, info)
        except Exception as e:  # the baseline is reporting only; never lose the GPU line to it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    ws_gb = (2 * w.nnz * 8 + (w.n_users + w.n_items) * eng.ld * 4 * 3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(w, name),
                       "sharding": ("ratings by nonzero along nnz-balanced user ranges; E_theta rows live with their owner, "
                                    "E_beta replicated") if world > 1 else "single GPU",
                       "combine": {"mc": ("item pass: per-rank row sums staged on the row's owner by the copy engines, added in rank order"
                                          if getattr(eng, "staged", False) else
                                          "item pass: per-rank row sums added in the NVSwitch (multimem.ld_reduce) by the row's owner")
                                         + f", Gamma update, new rows replicated with multimem.st; item rows in {eng.item_chunks} "
                                         "chunks, the exchange of a chunk overlaps the pass over the next",
                                   "nccl": "item pass: NCCL all-reduce of the row sums, every rank updates every row",
                                   "none": None}[eng.exchange],
                       "tiles": {"user_pass": len(eng.r.user_tiles), "item_pass": len(eng.r.item_tiles)},
                       "l2": (f"working set {ws_gb:.2f} GB (ratings + factor tables) vs 126 MB L2: "
                              + ("inputs larger than L2, no flush" if ws_gb > 0.5 else
                                 "comparable to L2 -- tables stay L2-resident between sweeps, as they do in a real fit; "
                                 "no flush (a flush would time a cold start no training loop sees)")),
                       "seg_len": {"user_pass": eng.r.seg_len_user, "item_pass": eng.r.seg_len_item}},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_fit_df": e2e_df, "parity_check": parity,
            "gpu_launches": (eng.launches_per_sweep + (len(eng.r.user_tiles) + 3 if with_elbo else 0)) * steps, "clocks": clocks}
    emit(line)
    job.finish()
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm: c1 (Gaussian MF), c4 (HPF-MAP epochs), topn -- one GPU
# ------------------------------------------------------------------------------------------------
def single_gpu_only(job):
    if job.rank != 0:
        job.finish()
        return False
    return True


def timed_steps(job, fn, steps, warmup):
    torch = job.torch
    job.preroll(fn)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run_c1(job):
    if not single_gpu_only(job):
        return 0
    import pandas as pd
    torch, args, steps, warmup = job.torch, job.args, job.steps, job.warmup
    from prob_matrix_factorization_b200.gaussian_mf_cavi_bias import GaussianMFCAVI, GaussianMFCAVIConfig
    job.world = 1
    w, u, i, x = make_workload("c1")
    mean = float(x.mean())
    xc = x.astype(np.float64) - mean                      # compare_models.py:54-58: the caller centres the ratings
    K, nnz = w.n_factors, w.nnz
    df = pd.DataFrame({"u": u.astype(np.int64), "i": i.astype(np.int64), "rating": xc})
    cfg = GaussianMFCAVIConfig(n_factors=K, max_iter=steps, tol=1e-3, random_state=42, verbose=False, **GAUSS_HP)
    m = GaussianMFCAVI(cfg, device=job.dev).fit(df, global_mean=mean)     # untimed first call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = GaussianMFCAVI(cfg, device=job.dev).fit(df, global_mean=mean)
    m_theta, m_beta = m.m_theta, m.m_beta                  # D2H of the factors (train_gaussian_full.py:77-80)
    dt = time.perf_counter() - t0
    e2e = {"value": nnz * steps / dt, "unit": UNIT, "h2d_bytes_per_step": 12 * nnz / steps,
           "d2h_bytes_per_step": (m_theta.nbytes + m_beta.nbytes) / 2 / steps, "seconds": dt, "sweeps": steps,
           "call": "GaussianMFCAVI(config).fit(DataFrame) -> m_theta / m_beta on the host (init draws included)"}
    eng = m._engine
    sampler = ClockSampler(job.my_gpus())
    sampler.start()
    sweep = lambda: eng.sweep(GAUSS_HP["sigma2"], GAUSS_HP["eta_theta2"], GAUSS_HP["eta_beta2"], GAUSS_HP["eta_bias2"])
    ms = timed_steps(job, sweep, steps, warmup)
    clocks = sampler.stop()
    tri = K * (K + 1) // 2
    alg = 2 * nnz * (4 * (tri + K) + 12) + (w.n_users + w.n_items) * 4 * (K * K + K) + 2 * nnz * (4 * K + 12) + (w.n_users + w.n_items) * 12
    peaks, peak_src = measured_peaks()
    achieved = alg / (ms / steps * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        v, info = cpu_port_run("c1", nnz, 3)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port",
               "sample": f"3 full sweeps of c1 by the oracle's C port (float64, OpenMP over rows), {info['seconds']:.2f}s"}
    emit({"metric": METRIC, "value": nnz * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
          "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic", "config": {"workload": describe(w, "c1"), "l2": "working set ~30 MB: L2-resident between sweeps, as "
                                          "in a real fit; no flush", "note": "float64 only inside the K x K Cholesky"},
          "roofline": {"bound": "hbm", "achieved": achieved, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                       "frac": achieved / float(peaks["hbm_gbs"]), "traffic": None, "peak_source": peak_src + " hbm_gbs",
                       "kernel": "pmf::gauss_* (4 passes per sweep)", "algorithmic_bytes_per_sweep": alg,
                       "note": "launch/latency-bound at this size (0.14 GB per sweep)"},
          "cpu_baseline": cpu, "e2e": e2e, "parity_check": None,
          "gpu_launches": getattr(eng, "launches_per_sweep", 8) * steps, "clocks": clocks})
    return 0


def run_c4(job):
    if not single_gpu_only(job):
        return 0
    torch, args, steps, warmup = job.torch, job.args, job.steps, job.warmup
    from prob_matrix_factorization_b200.hpf_pytorch import HPF_PyTorch, HPF_PyTorch_Config
    job.world = 1
    w, u, i, x = make_workload("c4")
    N, M, K, nnz, B = w.n_users, w.n_items, w.n_factors, w.nnz, 4096
    uc, ic = np.bincount(u, minlength=N), np.bincount(i, minlength=M)
    cfg = HPF_PyTorch_Config(n_factors=K, **MAP_HP)
    lazy = not args.dense_adam
    torch.manual_seed(0)
    m = HPF_PyTorch(N, M, uc, ic, cfg)
    m.fit_epochs(u, i, x, epochs=1, batch_size=B, lazy=lazy)           # untimed first call
    torch.cuda.synchronize()
    t0 = time.perf_counter()      # e2e: host arrays in, `steps` epochs, parameters back on the host
    losses = m.fit_epochs(u, i, x, epochs=steps, batch_size=B, lazy=lazy)
    theta = m.theta.detach().cpu().numpy(); beta = m.beta.detach().cpu().numpy()
    dt = time.perf_counter() - t0
    e2e = {"value": nnz * steps / dt, "unit": "ratings*epochs/s", "h2d_bytes_per_step": 16 * nnz / steps,
           "d2h_bytes_per_step": (theta.nbytes + beta.nbytes) / steps, "seconds": dt, "epochs": steps,
           "call": "HPF_PyTorch.fit_epochs(u, i, rating host arrays) -> theta / beta on the host"}
    sampler = ClockSampler(job.my_gpus())
    sampler.start()
    for _ in range(max(1, warmup // 2)):
        m.fit_epochs(u, i, x, epochs=1, batch_size=B, lazy=lazy)
    stats = {}
    m.fit_epochs(u, i, x, epochs=steps, batch_size=B, lazy=lazy, stats=stats)
    ms = float(stats["device_ms"])
    clocks = sampler.stop()
    peaks, peak_src = measured_peaks()
    alg = nnz * (2 * (K + 1) * 24 + 12)
    achieved = alg / (ms / steps * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        v, info = cpu_map_run(8)
        cpu = {"value": v, "unit": "ratings*epochs/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    emit({"metric": "rating-updates/sec (ratings*epochs/s)", "value": nnz * steps / (ms * 1e-3), "unit": "ratings*epochs/s",
          "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": describe(w, "c4") + f", batch {B}, one step = one epoch ({-(-nnz // B)} Adam steps)",
                     "adam": "lazy (touch-only, exactly equivalent to dense)" if lazy else "dense",
                     "l2": "parameters + moments 0.7 GB > L2; no flush"},
          "roofline": {"bound": "hbm", "achieved": achieved, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                       "frac": achieved / float(peaks["hbm_gbs"]), "traffic": None, "peak_source": peak_src + " hbm_gbs",
                       "kernel": "pmf::hpf_map_* (loss+grad+Adam per mini-batch)", "algorithmic_bytes_per_epoch": alg},
          "cpu_baseline": cpu, "e2e": e2e, "parity_check": None, "final_loss": float(losses[-1]),
          "gpu_launches": int(stats.get("launches", 0)), "clocks": clocks})
    return 0


def run_topn(job):
    if not single_gpu_only(job):
        return 0
    torch, args, steps, warmup = job.torch, job.args, job.steps, job.warmup
    from prob_matrix_factorization_b200 import _cabi
    from prob_matrix_factorization_b200.scoring import _as_table, top_n
    job.world = 1
    B, M, K, n = 8192, 230_000, 100, 50
    rng = np.random.default_rng(0)
    Fu_h = rng.gamma(0.3, 1.0, (B, K)).astype(np.float32)
    Fi_h = rng.gamma(0.3, 1.0, (M, K)).astype(np.float32)
    Fu, Fi = torch.from_numpy(Fu_h).to(job.dev), torch.from_numpy(Fi_h).to(job.dev)
    top_n(Fu[:256], Fi, n, tensor_cores=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, sc = top_n(Fu_h, Fi_h, n, tensor_cores=True)           # host factors in, NumPy indices / scores out
    dt = time.perf_counter() - t0
    e2e = {"value": B / dt, "unit": "user-rows/s", "h2d_bytes_per_step": Fu_h.nbytes + Fi_h.nbytes,
           "d2h_bytes_per_step": idx.nbytes + sc.nbytes, "seconds": dt, "call": "scoring.top_n(host factors) -> host indices/scores"}
    lib = _cabi.load()
    Tu, _ = _as_table(Fu, job.dev)
    Ti, _ = _as_table(Fi, job.dev)
    ld = Tu.shape[1]
    ws_bytes = lib.pmf_topn_workspace_bytes_ex(B, M, K, n, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=job.dev)
    d_idx = torch.empty((B, n), dtype=torch.int32, device=job.dev)
    d_sc = torch.empty((B, n), dtype=torch.float32, device=job.dev)
    call = lambda: _cabi.call("pmf_topn", Tu.data_ptr(), None, B, Ti.data_ptr(), M, K, ld, n, 1, d_idx.data_ptr(), d_sc.data_ptr(),
                              ws.data_ptr(), ws_bytes, None, _cabi.stream_ptr())
    sampler = ClockSampler(job.my_gpus())
    sampler.start()
    ms = timed_steps(job, call, steps, warmup)
    clocks = sampler.stop()
    peaks, peak_src = measured_peaks()
    flops = 2.0 * B * M * K
    achieved = flops / (ms / steps * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1500.0)))
    cpu = None
    if not args.no_cpu_baseline:
        v, info = cpu_topn_run(512)
        cpu = {"value": v, "unit": "user-rows/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    same = bool(np.array_equal(idx, d_idx.cpu().numpy()))
    emit({"metric": "user-rows/sec (dense U V^T top-50 over 230k items, K=100)", "value": B * steps / (ms * 1e-3), "unit": "user-rows/s",
          "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
          "vs_baseline": None, "dtype": "bf16 nomination + f32 exact re-score", "data": "synthetic",
          "config": {"workload": f"topn: {B} users x {M} items, K={K}, top-{n} (BASELINE.json configs[3], scoring half)",
                     "l2": "item table 92 MB (bf16 packed 52 MB) re-read per user tile pair; no flush"},
          "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                       "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained", "kernel": "pmf::topn_filter_kernel (tcgen05) + refine"},
          "cpu_baseline": cpu, "e2e": e2e, "parity_check": {"result": "ok" if same else "FAIL", "kind": "device-resident call == public API call (indices)"},
          "gpu_launches": None, "clocks": clocks})
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c3+elbo", "c4", "c5", "topn"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="ratings in the CPU-baseline sample")
    ap.add_argument("--seg-len", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fit-df", action="store_true", help="skip the fit(DataFrame) end-to-end leg")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--dense-adam", action="store_true", help="c4: dense Adam instead of the lazy (touch-only) one")
    ap.add_argument("--exchange", default=None, choices=["nccl", "mc", "ce"], help="multi-GPU combine of the item pass (default: PMF_EXCHANGE or the library's)")
    ap.add_argument("--tune", default="", help="comma list key=value passed to pmf_tune (experiments)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    job = Job(args)
    if args.workload == "c1":
        rc = run_c1(job)
    elif args.workload == "c4":
        rc = run_c4(job)
    elif args.workload == "topn":
        rc = run_topn(job)
    else:
        return run_cavi(job, args.workload)
    if job.rank == 0:
        real_world = int(os.environ.get("WORLD_SIZE", 1))
        if real_world > 1:
            job.world = real_world
            job.finish()
    return rc


if __name__ == "__main__":
    sys.exit(main())
