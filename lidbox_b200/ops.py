"""Thin Python wrappers over the TDNN entry points of the C-ABI (include/lidbox_b200.h).  Tensors are torch CUDA
tensors used purely as device memory; every function enqueues hand-written kernels on the current stream."""
import ctypes

import torch

from . import _lib

F32, BF16 = 0, 1

# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream:
# entries are (start_event, end_event, M, N, K, n_terms, layout)
GEMM_TIMER = None
# bench.py sets this to a list to record every GEMM descriptor of a step (replayed back-to-back for the roofline)
GEMM_RECORD = None


def replay(descs, device):
    """Re-issue recorded GEMM launches (single descriptors and grouped weight-gradient launches) on the current stream."""
    st = _lib.stream_ptr(device)
    lib = _lib.lib()
    for d, _keep, _shape in descs:
        if isinstance(d, tuple):
            _lib.check(lib.lbx_wgrad_grouped(d[0], d[1], st))
        else:
            _lib.check(lib.lbx_gemm_bf16(ctypes.byref(d), st))


def wgrad_grouped(problems, device):
    """lbx_wgrad_grouped: out += a^T . b for every problem (dicts with a, rows, a_cols, lda, a_off, b, b_cols, ldb, b_off,
    out, ldo, out_off; offsets in elements) in one persistent launch on the current stream."""
    if not problems:
        return
    arr = (_lib.WgradDesc * len(problems))()
    keep, shapes = [], []
    for d, q in zip(arr, problems):
        assert q["a"].dtype == torch.bfloat16 and q["b"].dtype == torch.bfloat16 and q["out"].dtype == torch.float32
        d.a, d.rows, d.a_cols, d.lda = _addr(q["a"], q.get("a_off", 0)), q["rows"], q["a_cols"], q["lda"]
        d.b, d.b_cols, d.ldb = _addr(q["b"], q.get("b_off", 0)), q["b_cols"], q["ldb"]
        d.out, d.ldo = _addr(q["out"], q.get("out_off", 0)), q["ldo"]
        keep.append((q["a"], q["b"], q["out"]))
        shapes.append((q["a_cols"], q["b_cols"], q["rows"]))
    if GEMM_RECORD is not None:
        GEMM_RECORD.append(((arr, len(problems)), keep, shapes))
    _lib.check(_lib.lib().lbx_wgrad_grouped(arr, len(problems), _lib.stream_ptr(device)))


def _addr(t, offset_elems=0):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def gemm(a, a_rows, a_cols, lda, b, b_rows, b_cols, ldb, out, ldo, *, layout=0, a_lo=None, b_lo=None, out_lo=None,
         a_off=0, b_off=0, out_off=0, k_splits=1, epi_atomic=False, bias=None, relu=False, rows_per_utt=0,
         valid_rows=0, mask_src=None, mask_off=0, accumulate=False, tile_n=0, colsum=None, colsum_off=0, colsum_mod=0,
         b1=None, b1_off=0, terms=None, b_map_rows=0, post_scale=None, post_shift=None):
    """lbx_gemm_bf16: see lbx_gemm_t.  `a`, `b` are bf16 buffers; offsets are in elements from their data_ptr."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    d = _lib.GemmDesc()
    d.a0, d.a1 = _addr(a, a_off), _addr(a_lo, a_off)
    d.a_rows, d.a_cols, d.lda = a_rows, a_cols, lda
    d.b0, d.b1 = _addr(b, b_off), (_addr(b_lo, b_off) if b1 is None else _addr(b1, b1_off))
    d.b_rows, d.b_cols, d.ldb = b_rows, b_cols, ldb
    d.layout = layout
    if terms is None:        # bf16x3 when residual planes are given: (a0,b0) + (a0,b1) + (a1,b0)
        terms = [(0, 0, 0), (0, 1, 0), (1, 0, 0)] if a_lo is not None else [(0, 0, 0)]
    d.n_terms = len(terms)
    for i, term in enumerate(terms):        # (A plane, B plane, A row offset[, B row offset[, output-column limit]])
        d.term_a[i], d.term_b[i], d.term_a_row[i] = term[0], term[1], term[2]
        d.term_b_row[i] = term[3] if len(term) > 3 else 0
        d.term_col_limit[i] = term[4] if len(term) > 4 else 0
    d.b_map_rows = b_map_rows
    d.post_scale, d.post_shift = _addr(post_scale), _addr(post_shift)
    d.k_splits = k_splits
    d.epi_atomic = int(epi_atomic)
    d.out_dtype = BF16 if out.dtype == torch.bfloat16 else F32
    assert out.dtype in (torch.bfloat16, torch.float32)
    d.out, d.out_lo, d.ldo = _addr(out, out_off), _addr(out_lo, out_off), ldo
    d.bias = _addr(bias)
    d.relu = int(relu)
    d.rows_per_utt, d.valid_rows = rows_per_utt, valid_rows
    d.mask_src = _addr(mask_src, mask_off)
    d.accumulate = int(accumulate)
    d.tile_n = tile_n
    d.colsum, d.colsum_mod = _addr(colsum, colsum_off), colsum_mod
    if GEMM_RECORD is not None:
        if layout == 0:
            shape = (a_rows, b_rows, a_cols)
        elif layout == 1:
            shape = (a_cols, b_cols, a_rows)
        else:
            shape = (a_rows, b_cols, a_cols)
        GEMM_RECORD.append((d, (a, b, out, a_lo, b_lo, out_lo, bias, mask_src, colsum), shape + (d.n_terms, layout)))
    if GEMM_TIMER is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().lbx_gemm_bf16(ctypes.byref(d), _lib.stream_ptr(a.device)))
        e1.record()
        if layout == 0:
            GEMM_TIMER.append((e0, e1, a_rows, b_rows, a_cols, d.n_terms, 0))
        elif layout == 1:
            GEMM_TIMER.append((e0, e1, a_cols, b_cols, a_rows, d.n_terms, 1))
        else:
            GEMM_TIMER.append((e0, e1, a_rows, b_cols, a_cols, d.n_terms, 2))
        return
    _lib.check(_lib.lib().lbx_gemm_bf16(ctypes.byref(d), _lib.stream_ptr(a.device)))
