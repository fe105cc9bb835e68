"""CPU tests (-m "not gpu"): the C-ABI library builds, loads and exports every symbol include/lidbox_b200.h declares;
host-only entry points (integer frame arithmetic, mel table) agree with the oracle.  No device call is made."""
import ctypes
import os
import re

import numpy as np

from oracle import lidbox_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lidbox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(built_lib):
    from lidbox_b200 import _lib
    handle = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(handle, name), "library does not export " + name
        assert name in _lib.SIGNATURES, "python binding missing for " + name
    assert sorted(_lib.SIGNATURES) == declared


def test_ms_to_frames_kat(built_lib):
    from lidbox_b200.features import audio
    for sr in range(1000, 60000, 1000):
        for ms in range(1, 5000, 100):
            assert audio.ms_to_frames(sr, ms) == (sr // 1000) * ms
    for sr in (8000, 11025, 22050, 44100, 48000):
        for ms in (1, 10, 25, 33):
            assert audio.ms_to_frames(sr, ms) == O.ms_to_frames(sr, ms)


def test_num_frames(built_lib):
    from lidbox_b200 import _lib
    lib = _lib.lib()
    for N in (0, 1, 399, 400, 401, 559, 560, 16000, 80000):
        for (L, s) in ((400, 160), (320, 160), (1, 1), (1600, 800)):
            assert lib.lbx_num_frames(N, L, s) == O.num_frames(N, L, s)


def test_mel_weight_matrix_matches_oracle(built_lib):
    from lidbox_b200.features import mel_ops
    for (m, k, sr, lo, hi) in [(40, 257, 16000, 0.0, 8000.0), (10, 129, 8000, 125.0, 3800.0),
                               (85, 513, 44100, 20.0, 11025.0), (25, 1025, 16000, 0.0, 8000.0)]:
        W = mel_ops.linear_to_mel_weight_matrix(m, k, sr, lo, hi).numpy()
        Wo = O.linear_to_mel_weight_matrix(m, k, sr, lo, hi)
        assert W.shape == Wo.shape
        # same fp32 evaluation order; logf (glibc) vs numpy log may differ in the last ulp
        np.testing.assert_allclose(W, Wo, atol=2e-5)
        assert ((W != 0) == (Wo != 0)).mean() > 0.999


def test_mel_band_packing(built_lib):
    from lidbox_b200 import _lib
    W = O.linear_to_mel_weight_matrix(40, 257, 16000, 0.0, 8000.0)
    start = np.empty(40, np.int32); length = np.empty(40, np.int32); off = np.empty(40, np.int32)
    packed = np.empty(W.size, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n = _lib.lib().lbx_mel_pack_bands(p(W), 257, 40, p(start), p(length), p(off), p(packed))
    assert n == 464
    R = np.zeros_like(W)
    for m in range(40):
        R[start[m]:start[m] + length[m], m] = packed[off[m]:off[m] + length[m]]
    assert np.array_equal(R, W)


def test_bad_arguments_raise_without_gpu(built_lib):
    from lidbox_b200 import _lib
    lib = _lib.lib()
    assert lib.lbx_spectrogram_f32(None, 1, 1000, 400, 160, 500, 2.0, None, None) == -2   # non power-of-two FFT
    assert b"fft_length" in lib.lbx_last_error()
    assert lib.lbx_spectrogram_f32(None, 1, 1000, 400, 160, 256, 2.0, None, None) == -2   # fft < frame
    assert lib.lbx_spectrogram_f32(None, -1, 1000, 400, 160, 512, 2.0, None, None) == -1
    assert lib.lbx_spectrogram_f32(None, 1, 100, 400, 160, 512, 2.0, None, None) == -1    # NULL signal with B*N > 0
    assert lib.lbx_spectrogram_f32(None, 0, 100, 400, 160, 512, 2.0, None, None) == 0     # empty batch is a no-op


def _plan(handle, shapes, workers, quad=0):
    from lidbox_b200 import _lib
    probs = (_lib.WgradDesc * len(shapes))()
    for d, (rows, a_cols, b_cols) in zip(probs, shapes):
        d.rows, d.a_cols, d.b_cols, d.lda, d.ldb, d.ldo = rows, a_cols, b_cols, a_cols, b_cols, b_cols
    seg = (ctypes.c_int * (6 * 4096))()
    n = ctypes.c_int(0)
    handle.lbx_wgrad_grouped_plan.restype = ctypes.c_int
    rc = handle.lbx_wgrad_grouped_plan(probs, len(shapes), workers, quad, seg, 4096, ctypes.byref(n))
    assert rc == 0
    return np.array(seg[:6 * n.value]).reshape(-1, 6)


def test_wgrad_grouped_partition_covers_every_k_block_once(built_lib):
    """lbx_wgrad_grouped cuts the (problem, tile, k-block) space of all weight-gradient problems into one contiguous
    range per CTA pair (stream-K).  The host-side replay runs the same decode function as the kernel: every k-block of
    every 256 x 256 tile must be covered exactly once, the ranges must be equal (+-1) and a worker must not see more
    than a handful of segments."""
    handle = ctypes.CDLL(built_lib)
    cases = [
        ([(26112, 1536, 512), (8704, 1536, 512), (8704, 512, 1500), (8704, 512, 512), (52224, 200, 512)], 74, 0),   # config 3
        ([(26112, 1536, 512), (8704, 1536, 512), (8704, 512, 1500), (8704, 512, 512), (52224, 200, 512)], 37, 1),   # 4-CTA clusters
        ([(5000, 512, 1500), (64, 64, 4), (3, 8, 8), (777, 264, 72), (12800, 1536, 512), (1, 512, 512)], 74, 0),
        ([(100, 128, 256)], 74, 0),                                                   # fewer k-blocks than workers
    ]
    for shapes, workers, quad in cases:
        seg = _plan(handle, shapes, workers, quad)
        live = [s for s in shapes if min(s) > 0]
        expect = {}
        for p, (rows, a_cols, b_cols) in enumerate(live):
            kb = -(-rows // 64)
            m_units = (-(-a_cols // 128) + 1) // 2
            n_tiles = -(-b_cols // 256) // (2 if quad else 1)
            for mu in range(m_units):
                for nt in range(n_tiles):
                    expect[(p, mu, nt)] = kb
        seen = {key: np.zeros(kb, dtype=np.int32) for key, kb in expect.items()}
        per_worker = {}
        for w, p, mu, nt, k0, k1 in seg:
            assert 0 <= k0 < k1 <= expect[(p, mu, nt)]
            seen[(p, mu, nt)][k0:k1] += 1
            per_worker.setdefault(w, []).append(k1 - k0)
        assert all((v == 1).all() for v in seen.values())                             # exactly once
        loads = [sum(v) for v in per_worker.values()]
        assert max(loads) - min(loads) <= 1                                           # equal shares
        assert max(len(v) for v in per_worker.values()) <= 6                          # few segments (few epilogues) per worker
