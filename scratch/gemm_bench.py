import sys, os
sys.path.insert(0, os.getcwd())
import torch
from lidbox_b200 import ops
dev = torch.device("cuda", 0)
bf = torch.bfloat16
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
def nt(name, M, N, K, lda=None, mask=False, acc=False, ldo=None, bias=True, relu=True, tile_n=0, out_f32=False):
    Kp = (K + 7) // 8 * 8
    lda = lda or Kp; ldo = ldo or (N + 7) // 8 * 8
    a = torch.randn((M * lda + 8 * Kp,), device=dev).to(bf)
    b = torch.randn((N, Kp), device=dev).to(bf)
    out = torch.zeros((M * ldo + 8 * N,), device=dev, dtype=torch.float32 if out_f32 else bf)
    msk = torch.randn((M * ldo + 8 * N,), device=dev).to(bf) if mask else None
    bs = torch.randn((N,), device=dev) if bias else None
    fn = lambda: ops.gemm(a, M, K, lda, b, N, K, Kp, out, ldo, bias=bs, relu=relu, mask_src=msk, accumulate=acc, tile_n=tile_n)
    us = t(fn)
    print("%-28s M=%6d N=%5d K=%5d tile_n=%3d  %8.1f us  %7.1f TFLOP/s" % (name, M, N, K, tile_n, us, 2.0 * M * N * K / us / 1e6))
def tn(name, Kc, M, N, lda=None, splits=1, tile_n=0):
    lda = lda or M
    ldb = (N + 7) // 8 * 8
    a = torch.randn((Kc * lda + 8 * M,), device=dev).to(bf)
    b = torch.randn((Kc, ldb), device=dev).to(bf)
    out = torch.zeros((M, N), device=dev)
    fn = lambda: ops.gemm(a, Kc, M, lda, b, Kc, N, ldb, out, N, layout=1, k_splits=splits, epi_atomic=True, tile_n=tile_n)
    us = t(fn)
    print("%-28s K=%6d M=%5d N=%5d splits=%2d tile_n=%3d %8.1f us  %7.1f TFLOP/s" % (name, Kc, M, N, splits, tile_n, us, 2.0 * M * N * Kc / us / 1e6))
B, P = 256, 34
only = sys.argv[1] if len(sys.argv) > 1 else None
if only:
    _nt, _tn = nt, tn
    nt = lambda name, *a, **k: _nt(name, *a, **k) if only in name else None
    tn = lambda name, *a, **k: _tn(name, *a, **k) if only in name else None
dense = len(sys.argv) > 2
for tile_n in ((64,) if dense else (256, 128)):
    nt("frame1 fwd", B * 6 * P, 512, 200, lda=40, tile_n=tile_n)
    nt("frame2 fwd", B * 3 * P, 512, 1536, lda=1024, tile_n=tile_n)
    nt("frame3 fwd", B * P, 512, 1536, tile_n=tile_n)
    nt("frame4 fwd", B * P, 512, 512, tile_n=tile_n)
    nt("frame5 fwd", B * P, 1500, 512, tile_n=tile_n)
    nt("frame5 dgrad", B * P, 512, 1500, mask=True, bias=False, relu=False, tile_n=tile_n)
    nt("frame4 dgrad", B * P, 512, 512, mask=True, bias=False, relu=False, tile_n=tile_n)
    nt("frame3 dgrad", B * P, 1536, 512, mask=True, bias=False, relu=False, tile_n=tile_n)
    nt("frame2 dgrad p1", B * 3 * P, 1024, 512, mask=True, bias=False, relu=False, ldo=1024, tile_n=tile_n)
    nt("frame2 dgrad p2 (acc)", B * 3 * P, 512, 512, mask=True, acc=True, bias=False, relu=False, ldo=1024, tile_n=tile_n)
    nt("segment1 fwd", B, 512, 3000, tile_n=tile_n)
    nt("segment2 fwd", B, 512, 512, tile_n=tile_n)
    nt("outputs fwd", B, 4, 512, tile_n=tile_n, out_f32=True, relu=False)
    nt("dH2 dgrad (K=4)", B, 512, 4, tile_n=tile_n, bias=False, relu=False, mask=True)
    nt("dH1 dgrad", B, 512, 512, tile_n=tile_n, bias=False, relu=False, mask=True)
    nt("segment1 dgrad", B, 3000, 512, bias=False, relu=False, out_f32=True, tile_n=tile_n)
    nt("big square", 8192, 8192, 8192, bias=False, relu=False, tile_n=tile_n)
    tn("frame5 wgrad", B * P, 512, 1500, splits=6, tile_n=tile_n)
    tn("frame4 wgrad", B * P, 512, 512, splits=18, tile_n=tile_n)
    tn("frame3 wgrad", B * P, 1536, 512, splits=6, tile_n=tile_n)
    tn("frame2 wgrad", B * 3 * P, 1536, 512, lda=1024, splits=6, tile_n=tile_n)
    tn("frame1 wgrad", B * 6 * P, 200, 512, lda=40, splits=37, tile_n=tile_n)
    tn("segment1 wgrad", B, 3000, 512, splits=3, tile_n=tile_n)
