"""Times the data-parallel optimizer kernel alone (torchrun, N GPUs): flag exchange + reduce-scatter + Adam + all-gather,
back to back, gradients untouched — the cost that sits between two training steps."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from lidbox_b200 import _lib
from lidbox_b200.models import xvector
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
m = xvector.create((198, 40), 4, precision="bf16", seed=0)
m.configure_optimizer(lr=1e-3)
m.enable_sharded_optimizer(dist.group.WORLD)
lib = _lib.lib()
sh = m._sharded
m.grads.normal_()
res = {}
def loop(n, with_wait):
    for _ in range(n):
        if with_wait:
            _lib.check(lib.lbx_dp_wait(_lib.ptr(sh["sig"]), sh["world"], _lib.ptr(sh["epoch"]), _lib.ptr(sh["local"]), _lib.stream_ptr(dev)))
        m._apply_sharded()
for with_wait in (False, True):
    loop(10, with_wait)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loop(20, with_wait)
    g.replay(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res["kernel_us_wait%d" % with_wait] = e0.elapsed_time(e1) * 1e3 / 100
m.dp_health()
st = sh["local"][4:10].view(torch.int64).cpu().tolist()
res["last_call_phase_us"] = {"barrier": (st[1] - st[0]) / 1e3, "reduce_adam_push": (st[2] - st[1]) / 1e3}
# single-GPU Adam for comparison
m1 = xvector.create((198, 40), 4, precision="bf16", seed=0)
m1.configure_optimizer(lr=1e-3)
m1.grads.normal_()
for _ in range(5):
    m1.apply_gradients()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    m1.apply_gradients()
e1.record()
torch.cuda.synchronize()
res["adam_1gpu_us"] = e0.elapsed_time(e1) * 1e3 / 100
res["nvls"] = bool(sh.get("mc_grads"))
if rank == 0:
    print(json.dumps(res))
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
