#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N-ce or $N-mc" > gpurun_out/j_pytest_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest_n$N.log; tail -8 gpurun_out/j_pytest_n$N.log | cut -c1-600
run_bench () {
  local tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline "$@" > gpurun_out/j_bench_$tag.json 2> gpurun_out/j_bench_$tag.log
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/j_bench_$tag.json'));r=d['roofline'];print('value %.3e ms %.3f (unpipelined %.3f) user %.3f item %.3f tiles %s parity %s' % (d['value'], d['ms_per_step'], (r.get('pipelined') or {}).get('unpipelined_ms_per_step', 0), r['user_pass_ms'], r['item_pass_ms'], d['config']['tiles'], (d.get('parity_check') or {}).get('result')))" 2>&1 | tail -1)"
}
PMF_ITEM_CHUNKS=1 PMF_EXCHANGE=ce run_bench ce_c1
PMF_ITEM_CHUNKS=4 PMF_USER_PASS_TILES=1 PMF_EXCHANGE=ce run_bench ce_c4shrink_u1
PMF_ITEM_CHUNKS=1 PMF_EXCHANGE=mc run_bench mc_c1 --no-parity
