#!/bin/bash
# round-2 final single-GPU validation: full suite, every bench workload, ncu evidence, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/l_pytest.log; tail -14 gpurun_out/l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; tail -2 gpurun_out/l_smoke.log
for wl in c5 c3+elbo c4 c1 c3 c2 topn; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/l_bench_$wl.json 2> gpurun_out/l_bench_$wl.log
  echo "bench $wl exit $? $(python -c "import json;d=json.load(open('gpurun_out/l_bench_$wl.json'));print('value %.3e ms/step %.4f frac %.3f e2e %s fit_df %s parity %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'] and round(d['e2e']['seconds']*1e3,1), (d.get('e2e_fit_df') or {}).get('seconds'), (d.get('parity_check') or {}).get('result')))" 2>&1 | tail -1)"
done
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.log; head -c 300 gpurun_out/l_bench_ref.json; echo
# ncu: launch lists (C1 sweep, C3+ELBO) and one full-set capture of a C5 sweep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 --launch-count 16 --csv --log-file gpurun_out/l_c1_launches.csv python scripts/exp/c1_short.py > gpurun_out/l_c1.log 2>&1; tail -1 gpurun_out/l_c1.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gamma_ --launch-skip 20 --launch-count 10 -o gpurun_out/l_c5_sweep python scripts/exp/c5_sweep.py c5 3 > gpurun_out/l_c5_ncu.log 2>&1; tail -2 gpurun_out/l_c5_ncu.log
ncu -i gpurun_out/l_c5_sweep.ncu-rep --page raw --csv > gpurun_out/l_c5_sweep_raw.csv 2>/dev/null
python scripts/ncu_traffic.py gpurun_out/l_c5_sweep_raw.csv c5/n1/tiles1x4 10 --out gpurun_out/l_dram_traffic.json
# compute-sanitizer on small shapes
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gamma.py -m gpu -x -q -k "tiled_passes_vs_oracle and 3-5-8 or coo_partition or count_keys or device_loop" > gpurun_out/l_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/l_sanitizer_memcheck.log; tail -4 gpurun_out/l_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_gauss.py tests/test_gpu_hpf_pytorch.py -m gpu -x -q -k "gaussian_vs_oracle and 10-64 or lazy_adam" > gpurun_out/l_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/l_sanitizer_racecheck.log; tail -4 gpurun_out/l_sanitizer_racecheck.log
