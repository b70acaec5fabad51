#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_hpf_pytorch.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/exp/c4_short.py > gpurun_out/h_c4.log 2>&1; cat gpurun_out/h_c4.log
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --launch-skip 700 --launch-count 2 -k regex:lazy_step --csv --log-file gpurun_out/h_c4_inst.csv python scripts/exp/c4_short.py > /dev/null 2>&1
cut -d, -f13- gpurun_out/h_c4_inst.csv | tail -4
