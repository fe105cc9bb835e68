#!/bin/bash
for f in 0 1 0 1; do
  LBX_GEMM_FAST_EPI=$f timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/ab_fast_$f.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_fast_$f.json").read())
print("fast=$f ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "gemm_ms", d["roofline"].get("kernel_ms_per_step"))
PY
done
