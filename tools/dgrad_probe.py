"""Isolated timing of the strided layer's one-launch data gradient (frame2 of config 3: rows 26112, C_in 512, C_out 512,
k 3, stride 2) with epilogue features switched off one by one — what bounds the launch?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidbox_b200 import ops
dev = torch.device("cuda")
rows, c, n_out, k, s = 26112, 512, 512, 3, 2
dz = (torch.randn(rows + 8, n_out, device=dev) * 0.01).bfloat16()
w = (torch.randn(k * c + 512, n_out, device=dev) * 0.03).bfloat16()
x = torch.relu(torch.randn(rows + 8, s * c, device=dev)).bfloat16()
out = torch.zeros(rows + 8, s * c, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(c, device=dev)
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)


def run(limit=True, mask=True, colsum=True, passes=2, ncols=None):
    lim = c if limit else 0
    terms = [(0, 0, 0, 0)] + ([(0, 0, -1, s * c, lim)] if passes == 2 else [])
    ops.gemm(dz, rows, n_out, n_out, w, ncols or s * c, n_out, n_out, out, s * c, b_map_rows=k * c, terms=terms,
             mask_src=x if mask else None, colsum=cs if colsum else None, colsum_mod=c)


def timeit(**kw):
    for _ in range(3):
        run(**kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(**kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {
    "full (2 passes, limit, mask, colsum)": timeit(),
    "no column limit": timeit(limit=False),
    "no colsum": timeit(colsum=False),
    "no mask, no colsum": timeit(mask=False, colsum=False),
    "1 pass, mask, colsum": timeit(passes=1),
    "1 pass, no mask, no colsum": timeit(passes=1, mask=False, colsum=False),
    "even half only (N=512), 2 passes, mask, colsum": timeit(ncols=c),
}
print(json.dumps(res, indent=1))
