#!/bin/bash
# GPU session recipe (run pieces through gpurun; every stage writes under gpurun_out/).  Written at the end of round 1,
# when the GPU budget ran out right after dgemm_dmma_fold was validated: stages 1-3 are what is still owed for it.
#
#   gpurun --timeout 300 -- 'bash tools/gpu_session.sh 1'
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2a}
case "${1:-help}" in
1)  # seconds: C-ABI checks of the folded kernel, of CPLX_NT and of the peer-store epilogue (P ranks emulated on one GPU)
    timeout 60 ./tools/fold_check --quick > gpurun_out/fold_quick.log 2>&1; echo "quick rc=$?"
    timeout 120 ./tools/fold_check --extra > gpurun_out/fold_extra.log 2>&1; echo "extra rc=$?"; tail -30 gpurun_out/fold_extra.log ;;
2)  # the whole GPU suite (xfail-guarded tests of the opt-in paths report XPASS when they work)
    timeout 1500 python -m pytest tests -m gpu -x -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log ;;
3)  # bench line, ncu launch list, full captures of the two dominant kernels -> python tools/make_profiles.py $TAG afterwards
    python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
    ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 6 -c 3 -f -o gpurun_out/prof_dgemm256_$TAG \
        python tools/profile_step.py legendre 256 > /dev/null 2>&1
    ncu --set full --clock-control none --import-source on -k regex:fft2_kernel -s 6 -c 6 -f -o gpurun_out/prof_fft2_cheb256_$TAG \
        python tools/profile_step.py chebyshev 256 > /dev/null 2>&1
    ls -la gpurun_out ;;
4)  # 2 GPUs (gpurun --gpus 2): NCCL exchange vs chunked overlap vs peer-store exchange, same 512^3 workload
    for mode in "" "JFX_SLAB_CHUNKS=4" "JFX_SLAB_FUSED_PACK=1" "JFX_SLAB_P2P=1"; do
      env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
          tools/check_slab_ranks.py 128 2>&1 | tail -4
      env $mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
          bench.py --gpus 2 --steps 10 --warmup 3 > "gpurun_out/scale2_${mode:-nccl}.json" 2> "gpurun_out/scale2_${mode:-nccl}.err"
      echo "mode=${mode:-nccl} rc=$?"; tail -c 400 "gpurun_out/scale2_${mode:-nccl}.json"; echo
    done ;;
*)  sed -n 2,8p "$0" ;;
esac
