"""GPU tests (-m gpu) of the fused dense head (lbx_head_fwd / lbx_head_bwd: the two segment layers of
lidbox/models/xvector.py:61-63, forward and backward, one persistent launch each) against fp64 math on the same
bf16-rounded operands."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    assert torch.cuda.is_available()
    from lidbox_b200 import _lib
    return _lib


def _r(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("B,K1,N1,N2", [(256, 3000, 512, 512), (37, 104, 72, 40), (300, 3000, 512, 512), (8, 64, 8, 8),
                                        (1, 3000, 512, 512)])
def test_head_fwd_bwd(lib, B, K1, N1, N2):
    L = lib.lib()
    st = lib.stream_ptr(torch.device("cuda"))
    ld1, ld2 = N1, N2
    pooled = _r((B, K1), 1).bfloat16()
    w1 = _r((K1, ld1), 2, K1 ** -0.5).bfloat16()
    w2 = _r((N1, ld2), 3, N1 ** -0.5).bfloat16()
    b1, b2 = _r((N1,), 4, 0.3), _r((N2,), 5, 0.3)
    h1 = torch.full((B, N1), float("nan"), device="cuda", dtype=torch.bfloat16)
    h2 = torch.full((B, N2), float("nan"), device="cuda", dtype=torch.bfloat16)
    scratch = torch.full((8, B, N1), float("nan"), device="cuda")     # workspace: contents on entry do not matter
    sync = torch.zeros(512, dtype=torch.int32, device="cuda")
    outs = []
    for _ in range(2):                       # twice: the barrier state must carry over, and the result is bit-reproducible
        lib.check(L.lbx_head_fwd(_p(pooled), B, K1, _p(w1), ld1, _p(b1), N1, _p(w2), ld2, _p(b2), N2, _p(h1), _p(h2),
                                 _p(scratch), scratch.numel(), _p(sync), st))
        outs.append((h1.clone(), h2.clone()))
    torch.cuda.synchronize()
    assert int(sync[2]) == 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    z1 = pooled.double() @ w1.double() + b1.double()
    r1 = torch.relu(z1)
    a1 = pooled.double().abs() @ w1.double().abs() + b1.double().abs()
    assert bool(((h1.double() - r1).abs() <= 2e-5 * a1 + 2.0 ** -8 * r1.abs() + 1e-30).all())     # fp32 sums + bf16 rounding
    z2 = h1.double() @ w2.double() + b2.double()                                                  # from the kernel's own h1
    r2 = torch.relu(z2)
    a2 = h1.double().abs() @ w2.double().abs() + b2.double().abs()
    assert bool(((h2.double() - r2).abs() <= 2e-5 * a2 + 2.0 ** -8 * r2.abs() + 1e-30).all())

    # ---- backward
    dh2 = (_r((B, N2), 6, 0.01) * (h2 > 0)).bfloat16()
    dh1 = torch.full((B, N1), float("nan"), device="cuda", dtype=torch.bfloat16)
    gpool = torch.full((B, K1), float("nan"), device="cuda")
    dw1, dw2, db1 = _r((K1, ld1), 7), _r((N1, ld2), 8), _r((N1,), 9)          # accumulated into
    dw1_0, dw2_0, db1_0 = dw1.clone(), dw2.clone(), db1.clone()
    lib.check(L.lbx_head_bwd(_p(dh2), _p(pooled), _p(h1), B, K1, N1, N2, _p(w1), ld1, _p(w2), ld2, _p(dh1), _p(gpool),
                             _p(dw1), _p(db1), _p(dw2), _p(sync), st))
    torch.cuda.synchronize()
    assert int(sync[2]) == 0
    g1 = (dh2.double() @ w2.double().T) * (h1 > 0)
    ag1 = dh2.double().abs() @ w2.double().abs().T
    assert bool(((dh1.double() - g1).abs() <= 2e-5 * ag1 + 2.0 ** -8 * g1.abs() + 1e-30).all())
    # column sums are taken before the bf16 rounding of dh1
    assert bool(((db1 - db1_0).double() - g1.sum(0)).abs().max() <= 1e-4 * ag1.sum(0).max() + 1e-6)
    ref_dw2 = h1.double().T @ dh2.double()
    assert bool(((dw2 - dw2_0).double() - ref_dw2).abs().max() <= 1e-5 * (h1.double().abs().T @ dh2.double().abs()).max() + 1e-6)
    ref_gp = dh1.double() @ w1.double().T                                      # from the kernel's own (rounded) dh1
    assert bool(((gpool.double() - ref_gp).abs() <= 2e-5 * (dh1.double().abs() @ w1.double().abs().T) + 1e-30).all())
    ref_dw1 = pooled.double().T @ dh1.double()
    assert bool(((dw1 - dw1_0).double() - ref_dw1).abs().max() <= 1e-5 * (pooled.double().abs().T @ dh1.double().abs()).max() + 1e-6)


def test_model_paths_agree(lib, monkeypatch):
    """The same training step with the fused head / grouped weight gradients and with one GEMM launch per layer:
    losses equal, gradients equal up to fp32 summation order."""
    from lidbox_b200.models import xvector
    rng = np.random.default_rng(0)
    x = rng.standard_normal((24, 61, 40)).astype(np.float32)
    y = np.arange(24) % 4
    grads, losses = [], []
    for fused in ("1", "0"):
        monkeypatch.setenv("LBX_HEAD_FUSED", fused)
        monkeypatch.setenv("LBX_WGRAD_GROUPED", fused)
        m = xvector.create((61, 40), 4, precision="bf16", seed=11)
        losses.append(m.loss_and_grads(x, y).clone())
        grads.append(m.grads.clone())
        m.head_health()
    assert float((losses[0] - losses[1]).abs().max()) < 2e-3
    for ly in m.layers:
        lo, hi = ly["w_off"], ly["b_off"] + ly["ldw"]
        den = float(grads[1][lo:hi].abs().max()) + 1e-30
        assert float((grads[0][lo:hi] - grads[1][lo:hi]).abs().max()) / den < 2e-2, ly["name"]
