#!/bin/bash
mkdir -p gpurun_out
PMF_TRACE=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/t_bench.json 2> gpurun_out/t_bench.log
grep "pmf trace\|e2e" gpurun_out/t_bench.log | cut -c1-140 | head -60
