#!/bin/bash
# multi-GPU check of the cross-pass pipeline ($1 GPUs)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N-mc or $N-nccl" > gpurun_out/j_pytest_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest_n$N.log; tail -8 gpurun_out/j_pytest_n$N.log | cut -c1-400
run_bench () {
  local n=$1 tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/j_bench_$tag.json 2> gpurun_out/j_bench_$tag.log
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/j_bench_$tag.json'));r=d['roofline'];print('value %.3e ms %.3f (unpipelined %.3f) user %.3f item %.3f e2e %s parity %s' % (d['value'], d['ms_per_step'], (r.get('pipelined') or {}).get('unpipelined_ms_per_step', 0), r['user_pass_ms'], r['item_pass_ms'], d['e2e'] and round(d['e2e']['seconds']*1e3,1), (d.get('parity_check') or {}).get('result')))" 2>&1 | tail -1)"
}
PMF_TRACE=1 run_bench $N n${N} --no-fit-df --no-cpu-baseline
grep "pmf trace" gpurun_out/j_bench_n${N}.log | sed -n 14,26p | cut -c1-120
for ch in 1 2 8; do PMF_ITEM_CHUNKS=$ch run_bench $N n${N}_chunks$ch --no-e2e --no-cpu-baseline --no-parity; done
