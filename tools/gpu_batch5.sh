#!/bin/bash
# round-2 batch 5: reversed tile order of the second of two plane-major passes (L2 reuse), A/B/A/B
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do for e in JFX_FFT_REVERSE=0 JFX_FFT_REVERSE=1; do
  env $e python tools/bench_cheb3.py --tag "cheb $e" --reps 40 2>&1 | tail -1
  env $e python tools/bench_cheb3.py --tag "four $e" --basis four --reps 40 2>&1 | tail -1
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/batch5.log 2>&1
cat gpurun_out/batch5.log
