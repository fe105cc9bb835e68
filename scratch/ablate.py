import sys, os
sys.path.insert(0, os.getcwd())
import torch
import bench
from lidbox_b200 import _lib
from lidbox_b200.models import xvector
class A: batch=256; seconds=2
os.environ["LBX_BENCH_GRAPH"] = "0"
wl = bench.XVectorTrainWorkload(A, 0, 1)
dev = torch.device("cuda", 0)
wl.setup(dev)
m = wl.model
def graph_time(body, iters=50):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    return bench._time_cuda(g.replay, iters) * 1e3
feats = wl._features()
bufs = m._buffers(256, wl.T, True)
def full(): m.train_step(wl._features(), wl.y)
def no_logmel(): m.train_step(wl.feats, wl.y)
def no_adam():
    m.loss_and_grads(wl._features(), wl.y); m.grads.zero_(); m._grads_clean = True
def fwd_only(): m._forward(wl.feats, bufs, True)
def logmel_only(): wl._features()
def adam_only(): m.apply_gradients()
res = {}
for name, fn in [("full", full), ("no_logmel", no_logmel), ("no_adam(+memset)", no_adam), ("fwd_only(pack+frames+pool+dense)", fwd_only), ("logmel_only", logmel_only), ("adam_only", adam_only)]:
    res[name] = graph_time(fn)
    print("%-40s %8.1f us" % (name, res[name]))
m.overlap_wgrad = False
print("%-40s %8.1f us" % ("full, wgrads on main stream", graph_time(full)))
m.overlap_wgrad = True
_lib.lib().lbx_set_pdl(0)
print("%-40s %8.1f us" % ("full, PDL off", graph_time(full)))
