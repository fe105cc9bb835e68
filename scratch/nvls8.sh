#!/bin/bash
N=${1:-8}
for nv in 1 0; do
  LBX_DP_NVLS=$nv timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 200 --warmup 10 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -1 > gpurun_out/nvls_${N}_$nv.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/nvls_${N}_$nv.json").read())
print("NVLS=$nv", d["ms_per_step"], d["value"], d["config"].get("dp_exchange")[:90])
PY
done
