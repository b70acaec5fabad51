#!/bin/bash
# round-2 GPU call A: single-GPU test suite + tile sweep at C5
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 600 python scripts/exp_tiles.py "1x1,1x4,1x8,1x16,2x8,3x8,2x16" > gpurun_out/a_tiles.log 2>&1
echo "tiles exit $?" >> gpurun_out/a_tiles.log
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_tiles.log
