#!/bin/bash
# 2-GPU check of the prioritised combine stream + unequal chunks, with e2e trace
N=${1:-2}
mkdir -p gpurun_out
run_bench () {
  local n=$1 tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/j_bench_$tag.json 2> gpurun_out/j_bench_$tag.log
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/j_bench_$tag.json'));r=d['roofline'];print('value %.3e ms %.3f user %.3f item %.3f e2e %s parity %s' % (d['value'], d['ms_per_step'], r['user_pass_ms'], r['item_pass_ms'], d['e2e'] and round(d['e2e']['seconds']*1e3,1), (d.get('parity_check') or {}).get('result')))" 2>&1 | tail -1)"
}
PMF_TRACE=1 run_bench $N n${N} --no-fit-df --no-cpu-baseline
grep "pmf trace" gpurun_out/j_bench_n${N}.log | sed -n 14,40p | cut -c1-120
for ch in 1 2 3 6; do PMF_ITEM_CHUNKS=$ch run_bench $N n${N}_chunks$ch --no-e2e --no-cpu-baseline --no-parity; done
