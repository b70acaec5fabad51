#!/bin/bash
# round-2 GPU call I (8 GPUs): headline scaling bench + parity at N=8 / N=4, chunk sweep, multi-GPU tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/i_gpus.txt
run_bench () {  # n tag extra...
  local n=$1 tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29640 \
    bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/i_bench_$tag.json 2> gpurun_out/i_bench_$tag.log
  echo "== $tag exit $? : $(python -c "import json;d=json.load(open('gpurun_out/i_bench_$tag.json'));r=d['roofline'];print('value %.3e ms %.3f user %.3f item %.3f e2e %s parity %s' % (d['value'], d['ms_per_step'], r['user_pass_ms'], r['item_pass_ms'], d['e2e'] and round(d['e2e']['seconds']*1e3,1), d.get('parity_check')))" 2>&1 | tail -1)"
}
PMF_TRACE=1 run_bench 8 n8 --no-fit-df
grep "pmf trace" gpurun_out/i_bench_n8.log | tail -14
run_bench 8 n8_nccl --exchange nccl --no-fit-df --no-cpu-baseline
for ch in 1 2 8; do PMF_ITEM_CHUNKS=$ch run_bench 8 n8_chunks$ch --no-e2e --no-cpu-baseline --no-parity; done
run_bench 4 n4 --no-fit-df --no-cpu-baseline
run_bench 2 n2 --no-fit-df --no-cpu-baseline
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8-mc or 4-mc or 8-nccl" > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log; tail -15 gpurun_out/i_pytest.log
