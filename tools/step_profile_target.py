"""Profiling target: eager training steps of BASELINE config 3 (batch 256 x 2 s) between cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
      --log-file gpurun_out/launches.csv python tools/step_profile_target.py [batch seconds loss]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lidbox_b200.features import audio
from lidbox_b200.models import xvector

B, sec = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 2)
loss = sys.argv[3] if len(sys.argv) > 3 else "xent"
n_out = 4 if loss == "xent" else 50
N = sec * 16000
T = 1 + (N - 400) // 160
x = torch.randn((B, N), device="cuda") * 0.1
y = torch.arange(B, device="cuda", dtype=torch.int32) % n_out
m = xvector.create((T, 40), n_out, precision="bf16", seed=0, head="log_softmax" if loss == "xent" else "none")
m.configure_optimizer()
sink = m.feature_sink(B, T, training=True)
kw = dict(loss=loss) if loss == "xent" else dict(loss="ap", ap_classes=n_out)


def step():
    feats = audio.logmelspectrograms(x, 16000, out=sink)
    return m.train_step(feats, y, **kw)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
m.head_health()
print("loss", float(step().mean()))
