"""Phase timing of the fused head kernels (CTA 0 globaltimer stamps) at batch 256, K1 = 3000, N1 = N2 = 512."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidbox_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
st = _lib.stream_ptr(dev)
B, K1, N1, N2 = 256, 3000, 512, 512
p = lambda t: ctypes.c_void_p(t.data_ptr())
pooled = torch.randn(B, K1, device=dev).bfloat16()
w1 = (torch.randn(K1, N1, device=dev) * 0.02).bfloat16(); w2 = (torch.randn(N1, N2, device=dev) * 0.04).bfloat16()
b1 = torch.zeros(N1, device=dev); b2 = torch.zeros(N2, device=dev)
h1 = torch.zeros(B, N1, device=dev, dtype=torch.bfloat16); h2 = torch.zeros(B, N2, device=dev, dtype=torch.bfloat16)
dh2 = (torch.randn(B, N2, device=dev) * 0.01).bfloat16(); dh1 = torch.zeros_like(h1)
scratch = torch.zeros(8, B, N1, device=dev); gpool = torch.zeros(B, K1, device=dev)
dw1 = torch.zeros(K1, N1, device=dev); dw2 = torch.zeros(N1, N2, device=dev); db1 = torch.zeros(N1, device=dev)
sync = torch.zeros(512, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def fwd():
    _lib.check(L.lbx_head_fwd(p(pooled), B, K1, p(w1), N1, p(b1), N1, p(w2), N2, p(b2), N2, p(h1), p(h2), p(scratch), scratch.numel(), p(sync), st))
def bwd():
    _lib.check(L.lbx_head_bwd(p(dh2), p(pooled), p(h1), B, K1, N1, N2, p(w1), N1, p(w2), N2, p(dh1), p(gpool), p(dw1), p(db1), p(dw2), p(sync), st))
res = {}
for name, fn, nst in (("fwd", fwd, 3), ("bwd", bwd, 2)):
    for cold in (0, 1):
        ph = []
        for _ in range(6):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            t = sync[4:4 + 2 * (2 + 2 * nst)].view(torch.int64).cpu().tolist()
            row = {"event_us": e0.elapsed_time(e1) * 1e3}
            for s in range(nst):
                row["step%d_work" % s] = (t[2 + 2 * s] - (t[0] if s == 0 else t[1 + 2 * s])) / 1e3
                if s + 1 < nst:
                    row["step%d_barrier" % s] = (t[3 + 2 * s] - t[2 + 2 * s]) / 1e3
            ph.append(row)
        res["%s_%s" % (name, "cold" if cold else "warm")] = ph[-1]
if hasattr(L, "lbx_head_profile") or True:
    try:
        buf = (ctypes.c_ulonglong * 8)()
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            L.lbx_head_profile(buf, 1)
            fn(); torch.cuda.synchronize()
            L.lbx_head_profile(buf, 1)
            res["prof_" + name] = {"wait_cycles_per_chunk": buf[0] / max(1, buf[2]), "compute_cycles_per_chunk": buf[1] / max(1, buf[2]), "chunks": buf[2], "items": buf[7], "prologue_cyc_per_item": buf[4] / max(1, buf[7]), "loop_cyc_per_item": buf[5] / max(1, buf[7]), "epilogue_cyc_per_item": buf[6] / max(1, buf[7])}
    except AttributeError:
        pass
print(json.dumps(res, indent=1))
