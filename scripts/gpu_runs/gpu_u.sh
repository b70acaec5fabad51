#!/bin/bash
mkdir -p gpurun_out
for sl in 64 128 256 512; do
  timeout 200 python bench.py --workload c3 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-parity --seg-len $sl > gpurun_out/u_c3_$sl.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/u_c3_$sl.json'));r=d['roofline'];print('seg_len $sl: ms %.4f user %.4f item %.4f launches %s' % (d['ms_per_step'], r['user_pass_ms'], r['item_pass_ms'], d['gpu_launches']))"
done
