#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2-" > gpurun_out/m_pytest_n2.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest_n2.log; tail -12 gpurun_out/m_pytest_n2.log | cut -c1-900
