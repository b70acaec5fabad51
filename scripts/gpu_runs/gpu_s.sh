#!/bin/bash
# final single-GPU check of the committed tree: full GPU suite + smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log; tail -4 gpurun_out/s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
