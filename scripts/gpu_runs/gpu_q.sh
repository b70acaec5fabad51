#!/bin/bash
# final 2-GPU check: multi-GPU parity test (in-switch combine) + the default bench command the driver runs at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2-mc" > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/q_pytest.log; tail -5 gpurun_out/q_pytest.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/q_bench_n2.json 2> gpurun_out/q_bench_n2.log
echo "bench exit $?"; grep "bench\] e2e" gpurun_out/q_bench_n2.log | sort -u | head -4
python -c "
import json;d=json.load(open('gpurun_out/q_bench_n2.json'));r=d['roofline']
print('value %.3e ms %.3f user %.3f item %.3f e2e %.1f ms fit_df %.2f s parity %s clocks %s' % (d['value'], d['ms_per_step'], r['user_pass_ms'], r['item_pass_ms'], d['e2e']['seconds']*1e3, d['e2e_fit_df']['seconds'], d['parity_check'], d['clocks']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/q_bench_ref_n2.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/q_bench_ref_n2.json'));print('reference arm at N=2: value %.3e cores %d' % (d['value'], d['cpu_baseline']['cores']))"
