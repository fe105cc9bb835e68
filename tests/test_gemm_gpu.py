"""GPU tests (-m gpu) of the tcgen05 GEMM behind the TDNN: compared with an fp64 matmul of the SAME bf16-rounded
operands (so the only difference is fp32 accumulation order): tolerance 1e-5 * sum|a||b| per element."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    assert torch.cuda.is_available()
    from lidbox_b200 import ops as o
    return o


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale)


def _check(out, ref, absref, tol=1e-5):
    err = (out.double() - ref).abs()
    bound = tol * absref + 1e-30
    assert bool((err <= bound).all()), "max err/bound = %g" % float((err / bound).max())


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (300, 512, 200), (1000, 1500, 512),
                                   (64, 4, 512), (7, 512, 3000), (20000, 512, 1536)])
def test_nt_plain(ops, M, N, K):
    a = _rand((M, K), 1).bfloat16()
    b = _rand((N, K), 2, 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, M, K, K, b, N, K, K, out, N)
    ref = a.double() @ b.double().T
    _check(out, ref, a.double().abs() @ b.double().abs().T)


def test_nt_bias_relu_bf16_out_and_lo(ops):
    M, N, K = 513, 1500, 512
    a = _rand((M, K), 3).bfloat16()
    b = _rand((N, K), 4, 0.05).bfloat16()
    bias = _rand((N,), 5)
    out = torch.zeros((M, 1504), device="cuda", dtype=torch.bfloat16)
    lo = torch.zeros_like(out)
    ops.gemm(a, M, K, K, b, N, K, K, out, 1504, bias=bias, relu=True, out_lo=lo)
    ref = torch.relu(a.double() @ b.double().T + bias.double())
    got = out[:, :N].double() + lo[:, :N].double()
    absref = a.double().abs() @ b.double().abs().T + bias.double().abs()
    _check(got, ref, absref, tol=2e-5)          # hi + lo carries ~16 mantissa bits
    assert float((out[:, :N].double() - ref).abs().max()) < 0.02 * float(ref.abs().max())
    assert bool((out[:, N:] == 0).all())


def test_nt_bf16x3_matches_fp32_inputs(ops):
    M, N, K = 400, 512, 1536
    a = _rand((M, K), 6)
    b = _rand((N, K), 7, 0.03)
    ah, bh = a.bfloat16(), b.bfloat16()
    al, bl = (a - ah.float()).bfloat16(), (b - bh.float()).bfloat16()
    out = torch.empty((M, N), device="cuda")
    ops.gemm(ah, M, K, K, bh, N, K, K, out, N, a_lo=al, b_lo=bl)
    ref = a.double() @ b.double().T
    absref = a.double().abs() @ b.double().abs().T
    _check(out, ref, absref, tol=3e-5)
    # normwise: what the fp32 forward config needs (<< 1e-4)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 2e-5
    # and the plain bf16 result is visibly worse, i.e. the extra terms are really accumulated
    out1 = torch.empty((M, N), device="cuda")
    ops.gemm(ah, M, K, K, bh, N, K, K, out1, N)
    assert float((out1.double() - ref).abs().max()) > 20 * float((out.double() - ref).abs().max())


def test_nt_overlapping_conv_view(ops):
    # causal strided Conv1D as a GEMM over an overlapping A view: k=3, stride=2, C=64, per-utterance pitch = 2*R
    B, C, k, s, R, T_out, N = 5, 64, 3, 2, 20, 18, 512
    Tpad = s * R
    x = torch.zeros((B * Tpad + 8, C), device="cuda", dtype=torch.bfloat16)
    data = _rand((B, 2 * T_out - 1, C), 8).bfloat16()
    xv = x[:B * Tpad].view(B, Tpad, C)
    xv[:, k - 1:k - 1 + data.shape[1]] = data
    w = _rand((N, k * C), 9, 0.05).bfloat16()
    out = torch.zeros((B * R + 2, N), device="cuda")
    ops.gemm(x, B * R, k * C, s * C, w, N, k * C, k * C, out, N, rows_per_utt=R, valid_rows=T_out, out_off=2 * N,
             relu=True)
    xp = xv.double()
    cols = torch.stack([xp[:, t * s:t * s + k].reshape(B, k * C) for t in range(T_out)], dim=1)
    ref = torch.relu(cols @ w.double().T)
    got = out[2:].view(B, R, N)
    _check(got[:, :T_out], ref, cols.abs() @ w.double().abs().T)
    assert bool((got[:, T_out:] == 0).all()) and bool((out[:2] == 0).all())


@pytest.mark.parametrize("Kc,M,N,splits", [(64, 128, 256, 1), (1000, 200, 512, 1), (5000, 1536, 512, 6),
                                           (2600, 512, 1500, 3), (256, 3000, 512, 2)])
def test_tn_wgrad_splitk_atomic(ops, Kc, M, N, splits):
    ldb = (N + 7) // 8 * 8                      # operand pitches are multiples of 8 elements (1500 -> 1504)
    a = _rand((Kc, M), 10).bfloat16()
    bfull = _rand((Kc, ldb), 11, 0.05).bfloat16()
    b = bfull[:, :N]
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(a, Kc, M, M, bfull, Kc, N, ldb, out, N, layout=1, k_splits=splits, epi_atomic=True)
    ref = a.double().T @ b.double()
    _check(out, ref, a.double().abs().T @ b.double().abs())
    # accumulates on top of what is there
    ops.gemm(a, Kc, M, M, bfull, Kc, N, ldb, out, N, layout=1, k_splits=splits, epi_atomic=True)
    _check(out, 2 * ref, 2 * (a.double().abs().T @ b.double().abs()))


def test_tn_overlapping_view_wgrad(ops):
    # weight gradient of the k=3, stride=2 conv: A^T view has pitch s*C < k*C
    B, C, k, s, R, N = 4, 64, 3, 2, 24, 512
    x = _rand((B * s * R + 8, C), 12).bfloat16()
    dy = _rand((B * R, N), 13).bfloat16()
    out = torch.zeros((k * C, N), device="cuda")
    ops.gemm(x, B * R, k * C, s * C, dy, B * R, N, N, out, N, layout=1, k_splits=2, epi_atomic=True)
    xf = x.double().reshape(-1)
    cols = torch.stack([xf[m * s * C:m * s * C + k * C] for m in range(B * R)])
    _check(out, cols.T @ dy.double(), cols.abs().T @ dy.double().abs())


def test_dgrad_mask_and_accumulate(ops):
    M, N, K = 300, 512, 512
    dz = _rand((M, K), 14).bfloat16()
    w = _rand((N, K), 15, 0.05).bfloat16()
    y = _rand((M, N), 16).bfloat16()
    out = _rand((M, N), 17).bfloat16()
    prev = out.clone()
    ops.gemm(dz, M, K, K, w, N, K, K, out, N, mask_src=y, accumulate=True)
    ref = torch.where(y.double() > 0, dz.double() @ w.double().T, torch.zeros((), device="cuda", dtype=torch.double))
    ref = ref + prev.double()
    assert float((out.double() - ref).abs().max()) < 0.01 * float(ref.abs().max()) + 0.05


@pytest.mark.parametrize("M,N,K", [(300, 512, 200), (1000, 1500, 512), (64, 4, 512), (5000, 512, 1536)])
def test_nn_forward_from_keras_layout(ops, M, N, K):
    # forward pass straight from the Keras-layout weight [K, N] (pitch padded to 8): B is an MN-major operand
    ldw = (N + 7) // 8 * 8
    a = _rand((M, K), 21).bfloat16()
    w = torch.zeros((K, ldw), device="cuda", dtype=torch.bfloat16)
    w[:, :N] = _rand((K, N), 22, 0.05).bfloat16()
    bias = _rand((N,), 23)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, M, K, K, w, K, N, ldw, out, N, layout=2, bias=bias)
    ref = a.double() @ w[:, :N].double() + bias.double()
    _check(out, ref, a.double().abs() @ w[:, :N].double().abs() + bias.double().abs())


def test_colsum_epilogue(ops):
    # bias gradient fused into the data-gradient epilogue: colsum[n % mod] += sum_m masked result
    M, N, K, C = 777, 1024, 512, 512
    dz = _rand((M, K), 24).bfloat16()
    w = _rand((N, K), 25, 0.05).bfloat16()
    y = _rand((M, N), 26).bfloat16()
    out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros((C,), device="cuda")
    ops.gemm(dz, M, K, K, w, N, K, K, out, N, mask_src=y, colsum=cs, colsum_mod=C)
    ref = torch.where(y.double() > 0, dz.double() @ w.double().T, torch.zeros((), device="cuda", dtype=torch.double))
    ref_cs = ref.sum(dim=0).view(2, C).sum(dim=0)
    assert float((cs.double() - ref_cs).abs().max()) < 1e-3 * float(ref.abs().sum(dim=0).max())


def _wgrad_problem(rows, a_cols, b_cols, seed, lda=None, ldo=None):
    """One weight-gradient problem: A [rows, a_cols] read through a view of pitch lda (lda < a_cols: overlapping rows,
    the implicit im2col of a Conv1D with kernel_size > strides), B [rows, b_cols], out fp32 [a_cols, ldo]."""
    lda = lda or a_cols
    ldo = ldo or -(-b_cols // 8) * 8
    flat = _rand(((rows - 1) * lda + a_cols + 64,), seed).bfloat16()
    ldb = -(-b_cols // 8) * 8
    b_full = _rand((rows, ldb), seed + 1, 0.05).bfloat16()
    b = b_full[:, :b_cols]
    out = _rand((a_cols, ldo), seed + 2)                       # the launch ACCUMULATES into whatever is there
    a_view = torch.as_strided(flat, (rows, a_cols), (lda, 1))
    ref = out[:, :b_cols].double() + a_view.double().T @ b.double()
    absref = out[:, :b_cols].double().abs() + a_view.double().abs().T @ b.double().abs()
    q = dict(a=flat, rows=rows, a_cols=a_cols, lda=lda, b=b_full, b_cols=b_cols, ldb=ldb, out=out, ldo=ldo)
    return q, ref, absref


@pytest.mark.parametrize("shapes", [
    [(1000, 128, 256)],                                        # one tile, one problem
    [(70, 200, 512, 40)],                                      # overlapping rows (frame1: k=5, C=40, stride 1), 2 k-blocks
    [(26112, 1536, 512), (8704, 1536, 512), (8704, 512, 1500), (8704, 512, 512), (52224, 200, 512, 40)],   # config 3
    [(5000, 512, 1500), (64, 64, 4), (3, 8, 8), (777, 264, 72), (12800, 1536, 512, 1024), (1, 512, 512)],
])
def test_wgrad_grouped(ops, shapes):
    """lbx_wgrad_grouped (all problems in one stream-K launch) vs fp64 A^T.B of the same bf16 operands; the pitch-padding
    columns of the outputs must stay untouched."""
    probs, refs = [], []
    for i, sh in enumerate(shapes):
        rows, a_cols, b_cols = sh[:3]
        q, ref, absref = _wgrad_problem(rows, a_cols, b_cols, 100 + 10 * i, lda=sh[3] if len(sh) > 3 else None)
        probs.append(q)
        refs.append((ref, absref, q["out"][:, b_cols:].clone()))
    ops.wgrad_grouped(probs, torch.device("cuda"))
    torch.cuda.synchronize()
    for q, (ref, absref, pad) in zip(probs, refs):
        _check(q["out"][:, :q["b_cols"]], ref, absref, tol=2e-5)
        assert torch.equal(q["out"][:, q["b_cols"]:], pad)


def test_wgrad_grouped_matches_per_layer_launches(ops):
    """Same problems through lbx_gemm_bf16 (layout TN, split-K, atomic epilogue): both paths sum the same products in
    fp32, only the order differs."""
    shapes = [(8704, 1536, 512), (8704, 512, 1500), (17408, 200, 512)]
    probs = []
    for i, (rows, a_cols, b_cols) in enumerate(shapes):
        q, _, _ = _wgrad_problem(rows, a_cols, b_cols, 300 + 10 * i, lda=40 if a_cols == 200 else None)
        q["out"].zero_()
        probs.append(q)
    singles = []
    for q in probs:
        o = torch.zeros_like(q["out"])
        ops.gemm(q["a"], q["rows"], q["a_cols"], q["lda"], q["b"], q["rows"], q["b_cols"], q["ldb"], o, q["ldo"],
                 layout=1, k_splits=9, epi_atomic=True)
        singles.append(o)
    ops.wgrad_grouped(probs, torch.device("cuda"))
    torch.cuda.synchronize()
    for q, o in zip(probs, singles):
        den = float(o.abs().max())
        assert float((q["out"] - o).abs().max()) < 2e-5 * den + 1e-6


def test_nt_passes_with_b_row_offsets_and_zero_fill_past_the_view(ops):
    """Accumulating passes with per-pass A / B row offsets: out[m, :] = sum_i A[m - i, :] . B[i*N : (i+1)*N, :]^T where
    rows of B past b_map_rows read as zeros — the one-GEMM gather form of a strided Conv1D's data gradient
    (k = 3 taps, stride 2: N = 2*C, pass 1 hangs over the end of the kernel)."""
    M, C, K = 700, 64, 128                       # C = C_in, K = C_out
    dz = _rand((M, K), 21).bfloat16()
    w = _rand((3 * C, K), 22, 0.05).bfloat16()   # Keras kernel rows [tap*C_in + c, C_out]
    junk = _rand((C, K), 23).bfloat16()          # memory behind the view: must NOT be read as data
    wbuf = torch.cat([w, junk]).contiguous()
    out = torch.full((M, 2 * C), float("nan"), device="cuda")
    ops.gemm(dz, M, K, K, wbuf, 2 * C, K, K, out, 2 * C, b_map_rows=3 * C, terms=[(0, 0, 0, 0), (0, 0, -1, 2 * C)])
    dzd, wd = dz.double(), w.double()
    prev = torch.cat([torch.zeros((1, K), device="cuda", dtype=torch.float64), dzd[:-1]])      # row m - 1 (zero before row 0)
    ref = torch.cat([dzd @ wd[:C].T + prev @ wd[2 * C:].T, dzd @ wd[C:2 * C].T], dim=1)
    absref = torch.cat([dzd.abs() @ wd[:C].abs().T + prev.abs() @ wd[2 * C:].abs().T, dzd.abs() @ wd[C:2 * C].abs().T], dim=1)
    _check(out, ref, absref)


@pytest.mark.parametrize("tile_n", [64, 128, 256])
def test_nn_forward_layout_all_tile_widths(ops, tile_n):
    """Forward layout (A K-major, B = Keras kernel [K, N] read as an MN-major operand) with bias + ReLU + bf16 output
    for every tile width the host may pick (256: CTA pairs; 128: small batches; 64: dense layers)."""
    M, N, K = 2176, 512, 1536
    a = _rand((M, K), 31).bfloat16()
    b = _rand((K, N), 32, 0.03).bfloat16()
    bias = _rand((N,), 33, 0.1)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, M, K, K, b, K, N, N, out, N, layout=2, bias=bias, relu=True, tile_n=tile_n)
    ref = torch.relu(a.double() @ b.double() + bias.double())
    absref = a.double().abs() @ b.double().abs() + bias.double().abs()
    assert bool(((out.double() - ref).abs() <= 2e-5 * absref + 2.0 ** -8 * ref.abs() + 1e-30).all())


def test_split_k_atomic_adds_the_bias_once(ops):
    """Split-K with the atomic epilogue into a zeroed fp32 output: every split adds its partial sums, the bias is added
    by the first split only (the embedding layer of the fp32 mode: 64 rows, K = 3000)."""
    M, N, K = 64, 512, 3000
    a = _rand((M, K), 41).bfloat16()
    b = _rand((K, N), 42, 0.02).bfloat16()
    bias = _rand((N,), 43)
    out = torch.zeros((M, N), device="cuda")
    ops.gemm(a, M, K, K, b, K, N, N, out, N, layout=2, bias=bias, tile_n=64, k_splits=11, epi_atomic=True)
    ref = a.double() @ b.double() + bias.double()
    _check(out, ref, a.double().abs() @ b.double().abs() + bias.double().abs(), tol=2e-5)


def test_nt_pass_column_limit_skips_zero_tiles(ops):
    """Same gather form with C_in = 256: the second pass only reaches the first 256 output columns
    (term_col_limit): tiles to the right skip it instead of multiplying by the zero fill — same result."""
    M, C, K = 900, 256, 192
    dz = _rand((M, K), 51).bfloat16()
    w = _rand((3 * C, K), 52, 0.05).bfloat16()
    outs = []
    for limit in (0, C):
        out = torch.full((M, 2 * C), float("nan"), device="cuda")
        ops.gemm(dz, M, K, K, w, 2 * C, K, K, out, 2 * C, b_map_rows=3 * C, terms=[(0, 0, 0, 0), (0, 0, -1, 2 * C, limit)])
        outs.append(out)
    dzd, wd = dz.double(), w.double()
    prev = torch.cat([torch.zeros((1, K), device="cuda", dtype=torch.float64), dzd[:-1]])
    ref = torch.cat([dzd @ wd[:C].T + prev @ wd[2 * C:].T, dzd @ wd[C:2 * C].T], dim=1)
    absref = torch.cat([dzd.abs() @ wd[:C].abs().T + prev.abs() @ wd[2 * C:].abs().T, dzd.abs() @ wd[C:2 * C].abs().T], dim=1)
    for out in outs:
        _check(out, ref, absref)
    assert torch.equal(outs[0], outs[1])
