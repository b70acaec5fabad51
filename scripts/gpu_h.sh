#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp/c4_short.py > gpurun_out/h_c4.log 2>&1; cat gpurun_out/h_c4.log
timeout 300 python -m pytest tests/test_gpu_hpf_pytorch.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.log; tail -2 gpurun_out/h_bench_c4.log; head -c 400 gpurun_out/h_bench_c4.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 --launch-count 12 --csv --log-file gpurun_out/h_c4_launches.csv python scripts/exp/c4_short.py > gpurun_out/h_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/h_c4_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
