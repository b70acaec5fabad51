#!/bin/bash
mkdir -p gpurun_out
PMF_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus 2 --steps 20 --warmup 5 --no-fit-df --no-cpu-baseline --no-parity > gpurun_out/r_bench_n2.json 2> gpurun_out/r_bench_n2.log
grep "pmf trace\|e2e" gpurun_out/r_bench_n2.log | cut -c1-140 | head -40
