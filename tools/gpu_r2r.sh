#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/diag_axes2.py 2>&1 | tail -5
python tools/diag_axes.py 256 4 2>&1 | grep -c "bad 0"
python tools/diag_in_nn.py 2>&1 | tail -10 | grep -v "\[(0, \[\]), (0, \[\]), (0, \[\]), (0, \[\])\]"
python tools/stress_first_call.py 512 4 2>&1 | tail -3
./tools/fold_check --quick 2>&1 | tail -1
./tools/fold_check --extra 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_at_size_gpu.py -q -s -k "not c4 and not fourier" 2>&1 | grep -E "single modes|\^3|passed|failed|Error|error" | tail
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.loads(open("gpurun_out/bench_r2r.json").read().splitlines()[-1])
print("value", b["value"], "ms/step", b["ms_per_step"], "frac", b["roofline"]["frac"], "detail", b["detail"], "fold_ab", b["fold_ab"]["legendre3_ms_per_pair_plain"], b["fold_ab"]["max_rel_diff_folded_vs_plain"])
PY
./tools/fold_check 2>&1 | grep -i " ms \|FOLD CHECK" | tail -8
