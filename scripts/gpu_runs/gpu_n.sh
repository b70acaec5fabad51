#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_c3elbo_launches.csv python scripts/exp/c3_elbo_short.py > gpurun_out/n_c3elbo.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/n_c3elbo_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tail=rows[-14:]
for r in tail: print(r[ki][:70], r[vi])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_c1_launches.csv python scripts/exp/c1_short.py > gpurun_out/n_c1.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/n_c1_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[-8:]: print(r[ki][:70], r[vi])
PY
