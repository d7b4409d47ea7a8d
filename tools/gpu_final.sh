#!/bin/bash
# end-of-round GPU session: the whole GPU suite, the bench line, the ncu launch list and the full captures that feed profiles/.
# The .ncu-rep files stay on the box (gpurun_out/ travels back only below 64 MiB): they are summarised there by
# tools/make_profiles.py into gpurun_out/profiles_$TAG/, which is copied into profiles/ afterwards.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2z}
W=/tmp/jfx_prof; mkdir -p $W
for e in JFX_NL_PARK=l2 JFX_NL_PARK=smem JFX_NL_PARK=l2; do echo "== $e"; env $e python tools/bench_nonlinear.py kdv 2>&1 | tail -1; env $e python tools/bench_nonlinear.py ch --n 1024 2>&1 | tail -1; done > gpurun_out/park_ab_$TAG.log 2>&1
cat gpurun_out/park_ab_$TAG.log
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
fi
python bench.py > $W/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cp $W/bench_$TAG.json gpurun_out/; tail -c 300 $W/bench_$TAG.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $W/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 6 -f -o $W/prof_fold256_$TAG \
    python tools/profile_step.py legendre 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6 -f -o $W/prof_fft2_cheb256_$TAG \
    python tools/profile_step.py chebyshev 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_rows -s 2 -c 1 -f -o $W/prof_fused_kdv_$TAG \
    python tools/bench_nonlinear.py kdv --batch 16384 > /dev/null 2>&1
JFX_PROF_IN=$W JFX_PROF_OUT=gpurun_out/profiles_$TAG python tools/make_profiles.py $TAG r2 > gpurun_out/make_profiles_$TAG.log 2>&1; tail -3 gpurun_out/make_profiles_$TAG.log
python tools/ncu_summary.py $W/prof_fused_kdv_$TAG.ncu-rep --ops > gpurun_out/profiles_$TAG/r2_fused_kdv_ncu_final.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
du -sh gpurun_out; ls gpurun_out gpurun_out/profiles_$TAG
