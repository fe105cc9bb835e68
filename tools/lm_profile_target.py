import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidbox_b200.features import audio
B, sec = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 5)
N = sec * 16000
T = 1 + (N - 400) // 160
x = torch.randn((B, N), device="cuda") * 0.1
out = torch.empty((B, T, 40), dtype=torch.float32, device="cuda")
for _ in range(4):
    audio.logmelspectrograms(x, 16000, out=out)
torch.cuda.synchronize()
