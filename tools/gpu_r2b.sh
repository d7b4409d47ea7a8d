#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_at_size_gpu.py tests/test_zz_fold_gpu.py -q -s -rxX > gpurun_out/r2b_atsize.log 2>&1; echo "atsize rc=$?"; grep -E "single modes|\^3|Cahn|KdV|passed|failed|Error|error" gpurun_out/r2b_atsize.log | tail -30
timeout 1500 python -m pytest tests -m gpu -x -q -rxX --deselect tests/test_at_size_gpu.py --deselect tests/test_zz_fold_gpu.py > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r2b_pytest_gpu.log
