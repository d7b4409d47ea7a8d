#!/bin/bash
# round-2 batch 4: per-pass (coalesced) twiddle tables + 64-thread fused CTAs: parity and timing
set -u
mkdir -p gpurun_out
run() { echo "=== $*"; env "$@" 2>&1 | grep -v "^$" | tail -8; }
{
run X=1 python -m pytest tests/test_fast_kernels_gpu.py tests/test_nonlinear_gpu.py tests/test_at_size_gpu.py tests/test_golden_large.py -x -q -m gpu
run X=1 python tools/bench_axes.py cheb
run X=1 python tools/bench_axes.py four2d --n 4096
run X=1 python tools/bench_axes.py batched --n 1024
run X=1 python tools/bench_nonlinear.py kdv
run X=1 python tools/bench_nonlinear.py ch --n 4096 --step
run X=1 python tools/bench_nonlinear.py ch --n 1024
run X=1 python tools/bench_cheb3.py --tag default
run X=1 python tools/bench_cheb3.py --tag four --basis four
} > gpurun_out/batch4.log 2>&1
tail -60 gpurun_out/batch4.log
