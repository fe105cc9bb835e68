"""Profiling target: BASELINE config 2 (fp32-grade embedding extraction, 64 x 2 s) between cudaProfilerStart/Stop."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidbox_b200.features import audio
from lidbox_b200.models import xvector
B, sec = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 2)
N = sec * 16000
T = 1 + (N - 400) // 160
x = torch.randn((B, N), device="cuda") * 0.1
m = xvector.create((T, 40), 4, precision="fp32", seed=0)
emb = xvector.as_embedding_extractor(m)
sink = m.feature_sink(B, T)
for _ in range(3):
    emb(audio.logmelspectrograms(x, 16000, out=sink))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    emb(audio.logmelspectrograms(x, 16000, out=sink))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
