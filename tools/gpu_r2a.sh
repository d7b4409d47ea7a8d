#!/bin/bash
# round-2 first GPU session: evidence owed for dgemm_dmma_fold + diagnosis of CPLX_NT
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
python - > gpurun_out/r2a_cplx.log 2>&1 <<'PY'
import os, subprocess, sys
sys.path.insert(0, "tests")
import importlib.util
spec = importlib.util.spec_from_file_location("t", "tests/test_zz_fold_gpu.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
e = dict(os.environ); e["JFX_CPLX_NT"] = "1"
r = subprocess.run([sys.executable, "-c", m._CPLX_SCRIPT % {"root": os.getcwd()}], capture_output=True, text=True, timeout=300, env=e)
print(r.stdout[-5000:]); print(r.stderr[-5000:]); print("rc", r.returncode)
PY
tail -5 gpurun_out/r2a_cplx.log
timeout 120 ./tools/fold_check --extra > gpurun_out/r2a_fold_extra.log 2>&1; echo "extra rc=$?"; tail -12 gpurun_out/r2a_fold_extra.log
timeout 600 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench_r2a.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6 -f -o gpurun_out/prof_fold256_r2a \
    python tools/profile_step.py legendre 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6 -f -o gpurun_out/prof_fft2_cheb256_r2a \
    python tools/profile_step.py chebyshev 256 > /dev/null 2>&1
ls -la gpurun_out
