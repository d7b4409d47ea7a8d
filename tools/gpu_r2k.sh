#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/fold_check --quick 2>&1 | tail -4
./tools/fold_check --extra 2>&1 | tail -3
python tools/diag_in_nn.py 2>&1 | tail -4
timeout 1200 python -m pytest tests/test_at_size_gpu.py -q -s -k "not c4 and not fourier" 2>&1 | grep -E "single modes|\^3|passed|failed|Error|error" | tail
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.loads(open("gpurun_out/bench_r2k.json").read().splitlines()[-1])
print("value", b["value"], "ms/step", b["ms_per_step"], "frac", b["roofline"]["frac"], "detail", b["detail"])
PY
./tools/fold_check 2>&1 | grep -i "ms\|FOLD CHECK" | tail -12
