#!/bin/bash
# A/B batch of the FFT / DCT axis-pass variants on one GPU (timing only; see tools/build_variant.py).  Build the variants first:
#   F=kernels_fft.cu,kernels_fft2.cu,kernels_fft2_pair.cu,kernels_fft2_stream.cu,kernels_fused.cu; O="--only=$F -DJFX_FFT2_ONLY=256"
#   python tools/build_variant.py p884    $O -DJFX_PLAN256_884 -DJFX_FFT_REGS8=64
#   python tools/build_variant.py p884r80 $O -DJFX_PLAN256_884
#   python tools/build_variant.py skip    $O -DJFX_FFT_SKIP_CORE
#   python tools/build_variant.py skip884 $O -DJFX_FFT_SKIP_CORE -DJFX_PLAN256_884 -DJFX_FFT_REGS8=64
# (recorded run: profiles/r2_fft_variants.txt)
set -u
mkdir -p gpurun_out
V=jaxfun_b200/variants
run() { echo "=== $*"; env "$@" 2>&1 | grep -v "^$" | tail -14; }
{
run JFX_TAG=base python tools/bench_axes.py cheb
run JFX_LIB_PATH=$V/libjfx_skip.so python tools/bench_axes.py cheb
run JFX_LIB_PATH=$V/libjfx_p884.so python tools/bench_axes.py cheb
run JFX_LIB_PATH=$V/libjfx_p884r80.so python tools/bench_axes.py cheb
run JFX_LIB_PATH=$V/libjfx_skip884.so python tools/bench_axes.py cheb
run JFX_FFT_STREAM=1 python tools/bench_axes.py cheb
run JFX_FFT_STREAM=2 python tools/bench_axes.py cheb
run JFX_FFT_STREAM=2 JFX_LIB_PATH=$V/libjfx_p884.so python tools/bench_axes.py cheb
run JFX_TAG=base python tools/bench_cheb3.py --tag base
run JFX_PAIR=1 python tools/bench_cheb3.py --tag pair
run JFX_SLAB_MB=48 python tools/bench_cheb3.py --tag slab48
run JFX_LIB_PATH=$V/libjfx_p884.so python tools/bench_cheb3.py --tag p884
run JFX_LIB_PATH=$V/libjfx_p884.so JFX_PAIR=1 python tools/bench_cheb3.py --tag p884+pair
run JFX_LIB_PATH=$V/libjfx_p884.so JFX_FFT_STREAM=2 python tools/bench_cheb3.py --tag p884+prefetch
run JFX_LIB_PATH=$V/libjfx_p884.so python -m pytest tests/test_fast_kernels_gpu.py -x -q -m gpu
run JFX_LIB_PATH=$V/libjfx_p884.so JFX_FFT_STREAM=2 python -m pytest tests/test_fast_kernels_gpu.py -x -q -m gpu
run JFX_FFT_STREAM=2 python -m pytest tests/test_fast_kernels_gpu.py -x -q -m gpu
} > gpurun_out/fft_batch1.log 2>&1
tail -5 gpurun_out/fft_batch1.log
{ timeout 300 python tools/slab_native_one_gpu.py 2; timeout 300 python tools/slab_native_one_gpu.py 4; } > gpurun_out/slab_native_one_gpu.log 2>&1
tail -12 gpurun_out/slab_native_one_gpu.log
