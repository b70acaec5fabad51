#!/bin/bash
# 8-GPU sweep of the combine pipeline variants
N=${1:-8}
mkdir -p gpurun_out
run_bench () {
  local tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29660 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/k_bench_$tag.json 2> gpurun_out/k_bench_$tag.log
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/k_bench_$tag.json'));r=d['roofline'];print('value %.3e ms %.3f (unpipelined %.3f) user %.3f item %.3f tiles %s' % (d['value'], d['ms_per_step'], (r.get('pipelined') or {}).get('unpipelined_ms_per_step', 0), r['user_pass_ms'], r['item_pass_ms'], d['config']['tiles']))" 2>&1 | tail -1)"
}
PMF_ITEM_CHUNKS=1 run_bench c1
PMF_ITEM_CHUNKS=4 PMF_CHUNK_SHAPE=shrink PMF_USER_PASS_TILES=1 run_bench c4shrink_u1
PMF_ITEM_CHUNKS=3 PMF_CHUNK_SHAPE=shrink PMF_USER_PASS_TILES=1 run_bench c3shrink_u1
PMF_ITEM_CHUNKS=2 PMF_CHUNK_SHAPE=equal run_bench c2equal_u2
PMF_ITEM_CHUNKS=2 PMF_CHUNK_SHAPE=shrink run_bench c2shrink_u2
PMF_ITEM_CHUNKS=4 PMF_CHUNK_SHAPE=equal run_bench c4equal_u4
