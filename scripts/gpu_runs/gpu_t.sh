#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gamma.py tests/test_gpu_fullsize.py -m gpu -x -q -k "golden or c2_full or c5_full_size or early" 2>&1 | tail -2
PMF_TRACE=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/t_bench.json 2> gpurun_out/t_bench.log
grep "e2e\|host init" gpurun_out/t_bench.log | cut -c1-140; grep "pmf trace" gpurun_out/t_bench.log | sed -n 26,38p | cut -c1-120
