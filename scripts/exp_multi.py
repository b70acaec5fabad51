"""Experiment driver (not part of the product): one torchrun launch, many (seg_len, tuning) combos on C5.
Usage: torchrun ... scripts/exp_multi.py "256:gamma_interleave=0" ":exchange=ce" ":exchange=mc" ...
(combo = [seg_len]:[key=value,...]; exchange = mc | ce | nccl; PMF_ITEM_CHUNKS / PMF_USER_PASS_TILES from the environment)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from prob_matrix_factorization_b200 import _cabi  # noqa: E402
from prob_matrix_factorization_b200.parallel import init_process_group  # noqa: E402


def main():
    rank, world, local = init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w, u, i, x = bench.make_workload(os.environ.get("EXP_WORKLOAD", "c5"))
    bench.pin(u); bench.pin(i); bench.pin(x)
    init = None
    for combo in sys.argv[1:]:
        seg, _, tune = combo.partition(":")
        for k in ("gamma_interleave", "gamma_group", "gamma_unroll"):
            _cabi.call("pmf_tune", k.encode(), 0)
        os.environ.pop("PMF_EXCHANGE", None)
        for kv in filter(None, tune.split(",")):
            k, v = kv.split("=")
            if k == "exchange":
                os.environ["PMF_EXCHANGE"] = v
            else:
                _cabi.call("pmf_tune", k.encode(), int(v))
        m = bench.make_model(w, 1, dev, (rank, world) if world > 1 else None, int(seg) if seg else None)
        if init is None:
            init = bench.initial_state_f32(m)
        m.fit_arrays(u, i, x, init)
        eng = m._engine
        for _ in range(5):
            eng.sweep()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        steps = 20
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for s in range(steps):
            ev[s][0].record(); eng.user_pass(); ev[s][1].record(); eng.item_pass(); ev[s][2].record()
        torch.cuda.synchronize()
        tu = np.mean([e[0].elapsed_time(e[1]) for e in ev]); ti = np.mean([e[1].elapsed_time(e[2]) for e in ev])
        tt = torch.tensor([tu, ti, ev[0][0].elapsed_time(ev[-1][2]) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"EXP [{eng.exchange}] {combo:40s} seg_len={eng.r.seg_len_user}/{eng.r.seg_len_item} user {tt[0]:.3f} item {tt[1]:.3f} step {tt[2]:.3f} ms "
                  f"-> {w.nnz / (tt[2].item() * 1e-3):.3e} nnz*it/s", flush=True)
        eng.close()
        m._engine = None
        del eng, m
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
