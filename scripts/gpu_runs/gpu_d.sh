#!/bin/bash
# round-2 GPU call D (2 GPUs): multi-GPU parity tests + bench at N=2 (mc / nccl) ; N is $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/d_gpus.txt
nvidia-smi topo -m >> gpurun_out/d_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/d_pytest_n$N.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest_n$N.log
tail -30 gpurun_out/d_pytest_n$N.log
for ex in mc nccl; do
  PMF_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 \
    bench.py --gpus $N --steps 20 --warmup 5 --exchange $ex --no-fit-df > gpurun_out/d_bench_n${N}_$ex.json 2> gpurun_out/d_bench_n${N}_$ex.log
  echo "bench $ex exit $?" >> gpurun_out/d_bench_n${N}_$ex.log
  tail -25 gpurun_out/d_bench_n${N}_$ex.log | cut -c1-300; head -c 1500 gpurun_out/d_bench_n${N}_$ex.json; echo
done
for ch in 1 2 8; do
  PMF_ITEM_CHUNKS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/d_bench_n${N}_chunks$ch.json 2> gpurun_out/d_bench_n${N}_chunks$ch.log
  echo "chunks $ch: $(python -c "import json;d=json.load(open('gpurun_out/d_bench_n${N}_chunks$ch.json'));print(d['ms_per_step'], d['roofline']['user_pass_ms'], d['roofline']['item_pass_ms'])")"
done
