#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gauss.py tests/test_gpu_hpf_extra.py tests/test_gpu_io.py -m gpu -x -q > gpurun_out/o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/o_pytest.log; tail -12 gpurun_out/o_pytest.log | cut -c1-300
for wl in c1 c3+elbo; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/o_bench_$wl.json 2> gpurun_out/o_bench_$wl.log
  echo "bench $wl exit $? $(python -c "import json;d=json.load(open('gpurun_out/o_bench_$wl.json'));print('value %.3e ms/step %.4f frac %.3f e2e %s elbo_ms %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'] and round(d['e2e']['seconds']*1e3,1), d['roofline'].get('elbo_ms')))" 2>&1 | tail -1)"
done
