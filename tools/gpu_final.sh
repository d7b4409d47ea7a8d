#!/bin/bash
# end-of-round GPU session: the whole GPU suite, the bench line, the ncu launch list and the full captures that feed profiles/
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2z}
for e in JFX_NL_PARK=smem JFX_NL_PARK=l2; do echo "== $e"; env $e python tools/bench_nonlinear.py kdv 2>&1 | tail -1; env $e python tools/bench_nonlinear.py ch --n 1024 2>&1 | tail -1; done > gpurun_out/park_ab_$TAG.log 2>&1
cat gpurun_out/park_ab_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$TAG.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6 -f -o gpurun_out/prof_fold256_$TAG \
    python tools/profile_step.py legendre 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6 -f -o gpurun_out/prof_fft2_cheb256_$TAG \
    python tools/profile_step.py chebyshev 256 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls gpurun_out | grep $TAG
