#!/bin/bash
N=${1:-2}
timeout 300 python -m pytest tests/test_dp_gpu.py -m gpu -q -s 2>&1 | grep -E "MEANDIFF|passed|failed|Error|error" | head -20
for nv in 1 0; do
  echo "== NVLS=$nv"
  LBX_DP_NVLS=$nv timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 200 --warmup 10 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -1 > gpurun_out/nvls_${N}_$nv.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/nvls_${N}_$nv.json").read())
print(d["ms_per_step"], d["value"], d["config"].get("dp_exchange"))
PY
done
