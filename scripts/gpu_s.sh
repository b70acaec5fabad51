#!/bin/bash
# final single-GPU check of the committed tree: full GPU suite, smoke, the driver's default bench command + reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log; tail -5 gpurun_out/s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/s_bench.json 2> gpurun_out/s_bench.log; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/s_bench.json'));r=d['roofline']
print('value %.4e ms %.3f frac %.3f traffic %s user %.3f item %.3f | e2e %.1f ms fit_df %.2f s | parity %s | launches %s | clocks %s %s' % (d['value'], d['ms_per_step'], r['frac'], r['traffic'], r['user_pass_ms'], r['item_pass_ms'], d['e2e']['seconds']*1e3, d['e2e_fit_df']['seconds'], d['parity_check']['result'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
timeout 200 python bench.py --impl reference > gpurun_out/s_bench_ref.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/s_bench_ref.json'));print('reference arm: value %.3e cores %d ms/step %.1f' % (d['value'], d['cpu_baseline']['cores'], d['ms_per_step']))"
