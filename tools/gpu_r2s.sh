#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -12 gpurun_out/r2s_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2s.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/bench_r2s.json").read().splitlines()[-1])
print("value", b["value"], "ms/step", b["ms_per_step"])
print("roofline", {k:b["roofline"][k] for k in ("achieved","frac","issued","avg_launch_ms")})
print("hbm", {k:b["roofline_hbm"][k] for k in ("achieved","frac")}, b["roofline_hbm"]["per_launch"]["frac"])
print("e2e", b.get("e2e",{}).get("value"), b.get("e2e",{}).get("ms_per_step"), "sync", b.get("e2e_sync",{}).get("ms_per_step"), "host_calls", b.get("e2e_host_calls",{}).get("ms_per_step"))
print("cpu", b.get("cpu_baseline",{}).get("value"))
print(json.dumps(b.get("legs"), indent=1)[:4000])
PY
