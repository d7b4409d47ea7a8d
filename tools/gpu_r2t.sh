#!/bin/bash
# multi-GPU session: correctness of the slab transform under every exchange mode, then the bench line of each
set -u
mkdir -p gpurun_out
NP=${NP:-2}
for mode in ${MODES:-JFX_SLAB_P2P=0 JFX_SLAB_P2P=1}; do
  tag=${mode}
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29533 \
      tools/check_slab_ranks.py 256 2>&1 | grep -E "SLAB CHECK|Error|error|route" | tail -12
  env $mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $NP --steps 10 --warmup 3 > "gpurun_out/scale${NP}_${tag}.json" 2> "gpurun_out/scale${NP}_${tag}.err"
  echo "mode=$tag rc=$?"
  python - "gpurun_out/scale${NP}_${tag}.json" <<'PY'
import json, sys
try:
    b = json.loads(open(sys.argv[1]).read().splitlines()[-1])
    print("   ms/step", round(b["ms_per_step"], 3), "eff", b.get("parallel_efficiency"), "single", b.get("single_gpu_same_size", {}).get("ms_per_step"), "check", b.get("check"), "e2e", b.get("e2e", {}).get("value"))
except Exception as e:
    print("   no line:", e)
PY
done
