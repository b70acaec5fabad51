#!/bin/bash
# round-2 GPU call E (1 GPU): conditional-graph probe, device loop + lazy Adam tests, c4 bench, tile sweep with per-launch seg_len
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -cudart static -o /tmp/condgraph_probe scripts/exp/condgraph_probe.cu && /tmp/condgraph_probe > gpurun_out/e_condgraph.log 2>&1
cat gpurun_out/e_condgraph.log
timeout 900 python -m pytest tests/test_gpu_gamma.py tests/test_gpu_hpf_pytorch.py tests/test_gpu_gauss.py tests/test_gpu_poisson_ext.py -m gpu -x -q > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log; tail -15 gpurun_out/e_pytest.log
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_c4.json 2> gpurun_out/e_bench_c4.log; tail -3 gpurun_out/e_bench_c4.log; head -c 700 gpurun_out/e_bench_c4.json; echo
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_c3.json 2> gpurun_out/e_bench_c3.log; tail -3 gpurun_out/e_bench_c3.log; head -c 400 gpurun_out/e_bench_c3.json; echo
timeout 600 python scripts/exp_tiles.py "1x4,1x8,1x16,2x8" > gpurun_out/e_tiles.log 2>&1; cat gpurun_out/e_tiles.log
