#!/bin/bash
mkdir -p gpurun_out
PMF_TRACE=1 timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline --no-fit-df > gpurun_out/g_c3_a.json 2> gpurun_out/g_c3_a.log
grep "device loop\|sweeps\|e2e" gpurun_out/g_c3_a.log | cut -c1-200
echo == steps 40
PMF_TRACE=1 timeout 300 python bench.py --workload c3 --steps 40 --warmup 3 --no-cpu-baseline --no-fit-df > gpurun_out/g_c3_b.json 2> gpurun_out/g_c3_b.log
grep "device loop\|sweeps\|e2e" gpurun_out/g_c3_b.log | cut -c1-200
echo == c2
PMF_TRACE=1 timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-fit-df > gpurun_out/g_c2.json 2> gpurun_out/g_c2.log
grep "device loop\|sweeps\|e2e" gpurun_out/g_c2.log | cut -c1-200
