#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2u.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-legs > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6 -f -o gpurun_out/prof_fold256_r2u \
    python tools/profile_step.py legendre 256 > /dev/null 2>&1
ls -la gpurun_out/*r2u*
