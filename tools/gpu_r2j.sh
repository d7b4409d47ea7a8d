#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/stress_first_call.py 512 6 2>&1 | tail -8
python tools/diag_in_nn.py 2>&1 | tail -10
timeout 1200 python -m pytest tests/test_at_size_gpu.py tests/test_zz_fold_gpu.py -q -s -rxX > gpurun_out/r2j_atsize.log 2>&1; echo "atsize rc=$?"; grep -E "single modes|\^3|Cahn|KdV|passed|failed|Error|error" gpurun_out/r2j_atsize.log | tail -30
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.loads(open("gpurun_out/bench_r2j.json").read().splitlines()[-1])
print("value", b["value"], "ms/step", b["ms_per_step"], "frac", b["roofline"]["frac"], "detail", b["detail"], "fold_ab", b.get("fold_ab"))
PY
./tools/fold_check --quick 2>&1 | tail -12
