#!/bin/bash
# round-2 batch 2: register-resident nonlinear kernel (parity + A/B timing) and the C-ABI slab transform on emulated ranks
set -u
mkdir -p gpurun_out
{
echo "=== parity"; timeout 900 python -m pytest tests/test_nonlinear_gpu.py tests/test_at_size_gpu.py -x -q -m gpu 2>&1 | tail -15
for e in JFX_NL_REG=1 JFX_NL_REG=0; do
  echo "=== $e"
  env $e python tools/bench_nonlinear.py kdv 2>&1 | tail -2
  env $e python tools/bench_nonlinear.py ch --n 4096 2>&1 | tail -2
  env $e python tools/bench_nonlinear.py ch --n 1024 2>&1 | tail -2
done
echo "=== slab native, emulated ranks"
timeout 300 python tools/slab_native_one_gpu.py 2 2>&1 | tail -8
timeout 300 python tools/slab_native_one_gpu.py 4 2>&1 | tail -8
} > gpurun_out/batch2.log 2>&1
tail -40 gpurun_out/batch2.log
