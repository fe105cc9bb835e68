import sys, os
sys.path.insert(0, os.getcwd())
import torch
os.environ["LBX_BENCH_GRAPH"] = "0"
import bench
class A: batch=256; seconds=2
wl = bench.XVectorTrainWorkload(A, 0, 1)
wl.setup(torch.device("cuda", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    wl.step()
torch.cuda.synchronize()
print("done")
