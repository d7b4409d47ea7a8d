#!/bin/bash
# build first:  python tools/build_variant.py lpb1 --only=kernels_fft2.cu -DJFX_FFT2_ONLY=4096 -DJFX_LPB4096=1
#               python tools/build_variant.py ft64 --only=kernels_fused.cu -DJFX_FUSED_THREADS=64   (64 is the default now)
# (recorded run: profiles/r2_fused_kdv_ncu.txt)
# round-2 batch 3: strided n = 4096 with one line per CTA (two CTAs per SM), 64-thread CTAs in the fused nonlinear kernel, ncu of the KdV kernel
set -u
mkdir -p gpurun_out
V=jaxfun_b200/variants
run() { echo "=== $*"; env "$@" 2>&1 | grep -v "^$" | tail -8; }
{
run JFX_TAG=base python tools/bench_axes.py four2d --n 4096
run JFX_LIB_PATH=$V/libjfx_lpb1.so python tools/bench_axes.py four2d --n 4096
run JFX_TAG=base python tools/bench_nonlinear.py ch --n 4096
run JFX_LIB_PATH=$V/libjfx_lpb1.so python tools/bench_nonlinear.py ch --n 4096
run JFX_LIB_PATH=$V/libjfx_lpb1.so python -m pytest tests/test_fast_kernels_gpu.py -x -q -m gpu -k 4096
run JFX_TAG=base python tools/bench_nonlinear.py kdv
run JFX_LIB_PATH=$V/libjfx_ft64.so python tools/bench_nonlinear.py kdv
run JFX_LIB_PATH=$V/libjfx_ft64.so python tools/bench_nonlinear.py ch --n 1024
run JFX_LIB_PATH=$V/libjfx_ft64.so python tools/bench_nonlinear.py ch --n 4096
run JFX_LIB_PATH=$V/libjfx_ft64.so python -m pytest tests/test_nonlinear_gpu.py -x -q -m gpu
} > gpurun_out/batch3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_rows -s 2 -c 1 -o gpurun_out/prof_fused_kdv_r2 -f python tools/bench_nonlinear.py kdv --batch 16384 > gpurun_out/prof_fused_kdv_r2.log 2>&1
tail -60 gpurun_out/batch3.log
