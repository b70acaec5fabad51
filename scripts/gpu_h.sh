#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hpf_pytorch.py tests/test_gpu_gamma.py tests/test_gpu_gauss.py -m gpu -x -q > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log; tail -12 gpurun_out/h_pytest.log
timeout 300 python scripts/exp/c4_short.py > gpurun_out/h_c4.log 2>&1; cat gpurun_out/h_c4.log
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.log; tail -2 gpurun_out/h_bench_c4.log; head -c 500 gpurun_out/h_bench_c4.json; echo
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_c3.json 2> gpurun_out/h_bench_c3.log; tail -3 gpurun_out/h_bench_c3.log
