#!/bin/bash
# round-2 GPU call F (1 GPU): device-loop timing probe, single-GPU suite, ncu launch list of a lazy-Adam epoch
mkdir -p gpurun_out
timeout 300 python scripts/exp/loop_timing.py > gpurun_out/f_loop_timing.log 2>&1; cat gpurun_out/f_loop_timing.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log; tail -25 gpurun_out/f_pytest.log
timeout 300 python scripts/exp/c4_short.py > gpurun_out/f_c4.log 2>&1; cat gpurun_out/f_c4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 24 --csv --log-file gpurun_out/f_c4_launches.csv python scripts/exp/c4_short.py > gpurun_out/f_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/f_c4_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:60], r[vi])
PY
