"""Warm graph-replay timings of the full training step and of the optimizer / pooling kernels alone."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import bench
from lidbox_b200 import _lib, ops
class A: batch=256; seconds=2
os.environ["LBX_BENCH_GRAPH"] = "0"
wl = bench.XVectorTrainWorkload(A, 0, 1)
dev = torch.device("cuda", 0)
wl.setup(dev)
m = wl.model
def graph_time(body, iters=100):
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
bufs = m._buffers(256, wl.T, True)
geo = bufs["geo"]; n = len(m.frames)
lib, st = _lib.lib(), _lib.stream_ptr(dev)
def full(): m.train_step(wl._features(), wl.y)
def no_logmel(): m.train_step(wl.feats, wl.y)
def adam_only(): m.apply_gradients()
def pool_fwd():
    _lib.check(lib.lbx_stats_pool_fwd(_lib.ptr(bufs["Y"]), ops.BF16, 256, geo.R[n - 1], geo.T[n], bufs["cn"], bufs["cnp"], 1e-10,
                                      _lib.ptr(bufs["pooled"]), _lib.ptr(bufs["var_raw"]), _lib.ptr(bufs["pooled_hi"]), None, st))
def pool_bwd():
    _lib.check(lib.lbx_stats_pool_bwd(_lib.ptr(bufs["Y"]), 256, geo.R[n - 1], geo.T[n], bufs["cn"], bufs["cnp"], 1e-10,
                                      _lib.ptr(bufs["pooled"]), _lib.ptr(bufs["var_raw"]), _lib.ptr(bufs["gpool"]),
                                      _lib.ptr(bufs["dZ"][n - 1]), ops._addr(m.grads, m.layers[n - 1]["b_off"]), 0, st))
m.train_step(wl.feats, wl.y)
out = [os.environ.get("LBX_LIB", "default").split("/")[-1]]
for name, fn in [("full", full), ("no_logmel", no_logmel), ("adam", adam_only), ("pool_fwd", pool_fwd), ("pool_bwd", pool_bwd)]:
    out.append("%s %.1f" % (name, graph_time(fn)))
print(" | ".join(out), flush=True)
