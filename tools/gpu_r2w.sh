#!/bin/bash
# evidence for profiles/: bench line, ncu launch list of the same command, full captures of the two dominant kernels
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2w}
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-legs > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6 -f -o gpurun_out/prof_fold256_$TAG \
    python tools/profile_step.py legendre 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6 -f -o gpurun_out/prof_fft2_cheb256_$TAG \
    python tools/profile_step.py chebyshev 256 > /dev/null 2>&1
ls -la gpurun_out/*$TAG*
