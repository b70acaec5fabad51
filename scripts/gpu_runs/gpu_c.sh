#!/bin/bash
# round-2 GPU call C (1 GPU): new tests + every bench workload through the rewritten bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gauss.py tests/test_gpu_gamma.py tests/test_gpu_hpf_pytorch.py -m gpu -x -q > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
for wl in c5 c3 c3+elbo c2 c1 c4 topn; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/c_bench_$wl.json 2> gpurun_out/c_bench_$wl.log
  echo "bench $wl exit $?" >> gpurun_out/c_bench_$wl.log
done
timeout 300 python bench.py --workload c4 --dense-adam --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c4_dense.json 2> gpurun_out/c_bench_c4_dense.log
timeout 300 python scripts/exp_tiles.py "1x2,1x3,1x4,1x5,1x4:gamma_chunk_reduce=1,1x4:gamma_chunk_reduce=0" > gpurun_out/c_tiles.log 2>&1
tail -3 gpurun_out/c_pytest.log; cat gpurun_out/c_tiles.log; for wl in c5 c3 c3+elbo c2 c1 c4 topn; do tail -2 gpurun_out/c_bench_$wl.log; head -c 600 gpurun_out/c_bench_$wl.json; echo; done
