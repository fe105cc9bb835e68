"""GPU parity tests (-m gpu): the CUDA feature kernels, called through the C-ABI, against the CPU oracle.

Tolerances (fp32 kernels vs the fp64 oracle):
  spectrograms / linear_to_mel : normwise per utterance, max|a-b| <= 1e-4 * max|b|  (SURVEY §7 hard part 6)
  log-mel                      : elementwise atol 1e-4 + rtol 1e-4 wherever the mel energy is >= 1e-4 (100x the
                                 log epsilon); atol 5e-3 below that, where the fp32 rounding floor of a 512-point FFT
                                 of a full-scale signal (~1e-9) is no longer negligible against the 1e-6 epsilon
  power_to_db                  : atol 2e-3 dB (20*log10 amplifies fp32 log rounding near the clip floor)
  frame counts / shapes        : exact
"""
import os

import numpy as np
import pytest
import torch

from oracle import lidbox_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def audio(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from lidbox_b200.features import audio as a
    return a


def _signals(B, N, seed=1234):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(N, dtype=torch.float64) / 16000.0
    f = 100.0 + 3900.0 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    x = 0.5 * torch.sin(2 * np.pi * f * t) + 0.05 * torch.randn(B, N, generator=g, dtype=torch.float64)
    return x.to(torch.float32).numpy()


def assert_logmel_close(out, ref, eps=1e-6):
    energy = np.exp(ref) - eps
    strong = energy >= 1e-4
    np.testing.assert_allclose(out[strong], ref[strong], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=5e-3)


def _normwise(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.float64)
    b = b.reshape(b.shape[0], -1).astype(np.float64)
    return (np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-30)).max()


def test_cfg1_sweeps_logmel(audio):
    # BASELINE config 1: 4 x 1 s @ 16 kHz exponential sine sweeps, amplitude -3 dBFS
    import scipy.signal
    t = np.arange(16000) / 16000.0
    sig = np.stack([0.7079 * scipy.signal.chirp(t, 100 * 2 ** i, 1.0, min(7900, 1600 * 2 ** i), method="logarithmic",
                                                phi=-90) for i in range(4)]).astype(np.float32)
    ref = O.logmel(sig, 16000, dtype=np.float64)
    out = audio.logmelspectrograms(sig, 16000).cpu().numpy()
    assert out.shape == (4, 98, 40)
    assert_logmel_close(out, ref)
    # unfused chain through the drop-in names gives the same result
    S = audio.spectrograms(sig, 16000)
    M = audio.linear_to_mel(S, 16000)
    assert_logmel_close(torch.log(M + 1e-6).cpu().numpy(), ref)


@pytest.mark.parametrize("B,sec", [(1, 1), (3, 2), (64, 2), (5, 5)])
def test_spectrogram_and_mel_values(audio, B, sec):
    sig = _signals(B, 16000 * sec)
    S_ref = O.spectrograms(sig, 16000, dtype=np.float64)
    S = audio.spectrograms(sig, 16000).cpu().numpy()
    assert S.shape == S_ref.shape == (B, 1 + (16000 * sec - 400) // 160, 257)
    assert _normwise(S, S_ref) < 1e-4
    M_ref = O.linear_to_mel(S_ref, 16000, dtype=np.float64)
    M = audio.linear_to_mel(S, 16000).cpu().numpy()
    assert _normwise(M, M_ref) < 1e-4
    lm = audio.logmelspectrograms(sig, 16000).cpu().numpy()
    assert_logmel_close(lm, O.log_eps(M_ref))


def test_golden_wav_fixtures(audio):
    g = np.load(os.path.join(GOLDEN, "wav_fixtures.npz"))
    sig = g["pcm"].astype(np.float32) / np.float32(32768.0)
    lm = audio.logmelspectrograms(sig, 16000).cpu().numpy()
    assert_logmel_close(lm, g["logmel"].astype(np.float64))
    S = audio.spectrograms(sig, 16000)
    np.testing.assert_allclose(S.sum(dim=2).cpu().numpy(), g["spec_rowsum"], rtol=1e-4)
    db = audio.power_to_db(S).cpu().numpy()[:, ::8, ::16]
    np.testing.assert_allclose(db, g["db"], atol=2e-3)


def test_reference_test_spectrograms_grid(audio):
    # /root/reference/tests/test_features_audio.py:131-145 on a synthetic 3 s signal + value parity (generic FFT path)
    s = _signals(1, 48000, seed=5)
    for len_ms in range(20, 101, 20):
        for n_fft in (256, 512, 1024, 2048):
            if n_fft < audio.ms_to_frames(16000, len_ms):
                continue
            step_ms = len_ms // 2
            P = audio.spectrograms(s, 16000, frame_length_ms=len_ms, frame_step_ms=step_ms, fft_length=n_fft)[0]
            P = P.cpu().numpy()
            assert not np.isnan(P).any()
            assert P.shape[0] == s.shape[1] // audio.ms_to_frames(16000, step_ms) - 1
            assert P.shape[1] == n_fft // 2 + 1
            ref = O.spectrograms(s, 16000, len_ms, step_ms, 2.0, n_fft, dtype=np.float64)[0]
            assert _normwise(P[None], ref[None]) < 1e-4


def test_reference_test_linear_to_mel_grid(audio):
    # tests/test_features_audio.py:147-155 + values
    s = _signals(1, 48000, seed=6)
    P = audio.spectrograms(s, 16000)
    P_ref = O.spectrograms(s, 16000, dtype=np.float64)
    for num_mel_bins in range(10, 100, 15):
        M = audio.linear_to_mel(P, 16000, num_mel_bins=num_mel_bins)[0].cpu().numpy()
        assert not np.isnan(M).any()
        assert M.shape == (P.shape[1], num_mel_bins)
        ref = O.linear_to_mel(P_ref, 16000, num_mel_bins=num_mel_bins, dtype=np.float64)[0]
        assert _normwise(M[None], ref[None]) < 1e-4
        # fused kernel with a non-default mel count
        lm = audio.logmelspectrograms(s, 16000, num_mel_bins=num_mel_bins)[0].cpu().numpy()
        assert_logmel_close(lm, np.log(ref + 1e-6))


def test_reference_test_power_to_db(audio):
    # tests/test_features_audio.py:115-123 (+ oracle values)
    s = _signals(2, 16000, seed=7)
    P = O.spectrograms(s, 16000)
    for top_db in range(10, 110, 10):
        db = audio.power_to_db(P, top_db=float(top_db)).cpu().numpy()
        assert not np.isnan(db).any()
        assert db.max() <= 0
        np.testing.assert_allclose(db, O.power_to_db(P, top_db=float(top_db)), atol=2e-3)


def test_power_and_frame_parameter_variants(audio):
    s = _signals(2, 12345, seed=8)        # odd length
    for (len_ms, step_ms, power, n_fft) in ((25, 10, 1.0, 512), (25, 10, 1.5, 512), (20, 7, 2.0, 512),
                                           (32, 16, 2.0, 512), (25, 10, 2.0, 1024), (5, 3, 2.0, 128)):
        ref = O.spectrograms(s, 16000, len_ms, step_ms, power, n_fft, dtype=np.float64)
        out = audio.spectrograms(s, 16000, len_ms, step_ms, power, n_fft).cpu().numpy()
        assert out.shape == ref.shape
        assert _normwise(out, ref) < 1e-4
    # 11.025 kHz: frame_length 275 (odd), frame_step 110 -> exercises the unaligned shared-memory path
    ref = O.spectrograms(s, 11025, dtype=np.float64)
    out = audio.spectrograms(s, 11025).cpu().numpy()
    assert out.shape == ref.shape and _normwise(out, ref) < 1e-4
    ref = O.spectrograms(s, 22050, 5, 1, dtype=np.float64)    # L = 110, step = 22
    out = audio.spectrograms(s, 22050, 5, 1).cpu().numpy()
    assert out.shape == ref.shape and _normwise(out, ref) < 1e-4


def test_edge_cases(audio):
    # shorter than one frame -> T == 0; exactly one frame; ragged tail; empty batch
    assert audio.spectrograms(np.zeros((2, 399), np.float32), 16000).shape == (2, 0, 257)
    assert audio.logmelspectrograms(np.zeros((2, 399), np.float32), 16000).shape == (2, 0, 40)
    assert audio.spectrograms(np.zeros((0, 16000), np.float32), 16000).shape == (0, 98, 257)
    one = _signals(1, 400, seed=9)
    assert_logmel_close(audio.logmelspectrograms(one, 16000).cpu().numpy(), O.logmel(one, 16000, dtype=np.float64))
    for N in (559, 560, 561, 400 + 160 * 32, 400 + 160 * 33 - 1):
        x = _signals(2, N, seed=N)
        ref = O.logmel(x, 16000, dtype=np.float64)
        out = audio.logmelspectrograms(x, 16000).cpu().numpy()
        assert out.shape == ref.shape
        assert_logmel_close(out, ref)
    # all-zero signal: log(0 + 1e-6)
    z = audio.logmelspectrograms(np.zeros((1, 16000), np.float32), 16000).cpu().numpy()
    np.testing.assert_allclose(z, np.log(1e-6), rtol=1e-6)
    with pytest.raises(ValueError):
        audio.spectrograms(np.zeros(16000, np.float32), 16000)          # rank 1 (tf_utils.py:168)
    with pytest.raises(NotImplementedError):
        audio.spectrograms(np.zeros((1, 16000), np.float32), 16000, fft_length=500)


def test_full_size_properties(audio):
    # BASELINE sweep maximum (2048 x 5 s): linearity and batch-invariance instead of a CPU comparison
    B, N = 2048, 80000
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, N, generator=g, device="cuda") * 0.1
    S = audio.spectrograms(x[:64], 16000)
    S2 = audio.spectrograms(2.0 * x[:64], 16000)
    assert torch.allclose(S2, 4.0 * S, rtol=1e-5, atol=0)              # |2x|^2 = 4|x|^2 exactly in fp32
    lm = audio.logmelspectrograms(x, 16000)
    assert lm.shape == (B, 498, 40) and torch.isfinite(lm).all()
    # each utterance is independent of its position in the batch
    lm_b = audio.logmelspectrograms(x[1000:1003], 16000)
    assert torch.equal(lm_b, lm[1000:1003])
    # Parseval on the power spectrogram of one frame: sum_k c_k |X_k|^2 = 512 * sum_n (x_n w_n)^2
    fr = x[7, :400].double().cpu().numpy() * O.hann_window(400, np.float64)
    P = S[7, 0].double().cpu().numpy()
    lhs = P[0] + P[256] + 2 * P[1:256].sum()
    assert abs(lhs - 512 * (fr ** 2).sum()) / lhs < 1e-5


def test_feature_normalisation_parity(built_lib):
    # lidbox/features/__init__.py: fp32 kernels vs the fp64 oracle, tolerance 1e-4 normwise
    from lidbox_b200 import features as F
    rng = np.random.default_rng(20)
    for shape in ((4, 198, 40), (1, 1, 1), (3, 17, 5), (2, 498, 40)):
        x = (rng.standard_normal(shape) * 3 - 5).astype(np.float32)
        for axis in range(3):
            assert _normwise(F.cmn(x, axis=axis).cpu().numpy(), O.cmn(x, axis)) < 1e-4 or x.shape[axis] == 1
            # (x - mean) / std is ill-conditioned in fp32 when std << |x| (e.g. two nearly equal values): the bound is
            # 1e-4 plus the fp32 rounding of x relative to std
            ref = O.cmvn(x, axis)
            std = x.astype(np.float64).std(axis=axis, keepdims=True)
            tol = 1e-4 + 4e-7 * np.abs(x).max() / np.maximum(std, 1e-30)
            assert (np.abs(F.cmvn(x, axis=axis).cpu().numpy() - ref) <= tol * (1 + np.abs(ref))).all()
        for axis in (None, 0, 1, 2):
            y = F.feature_scaling(x, -1.0, 2.5, axis=axis).cpu().numpy()
            np.testing.assert_allclose(y, O.feature_scaling(x, -1.0, 2.5, axis=axis), rtol=1e-5, atol=1e-5)
    # reference property tests (tests/test_features.py:28-58) on the CUDA path
    for _ in range(5):
        x = rng.uniform(-100, 100, size=rng.integers(1, 20, size=3)).astype(np.float32)
        for axis in range(3):
            y_mv = F.cmvn(x, axis=axis).cpu().numpy()
            assert not np.isnan(y_mv).any() and y_mv.shape == x.shape
            assert np.abs(y_mv.mean(axis=axis)).max() < 0.1 and y_mv.var(axis=axis).max() < 10
        for window_len in [-1] + list(range(2, x.shape[0] + 1)):
            for nv in (True, False):
                y = F.window_normalization(x, axis=1, window_len=window_len, normalize_variance=nv).cpu().numpy()
                assert not np.isnan(y).any() and y.shape == x.shape
                if not nv:
                    np.testing.assert_allclose(y, O.window_normalization(x, 1, window_len, nv), rtol=2e-4, atol=2e-3)
                elif (x.shape[1] if window_len == -1 else min(window_len, x.shape[1])) >= 5:
                    # dividing by the std of a 2-4 sample window is ill-conditioned in fp32 (nearly equal samples)
                    np.testing.assert_allclose(y, O.window_normalization(x, 1, window_len, nv), rtol=2e-3, atol=2e-3)
    x = (rng.standard_normal((8, 298, 40)) * 2 + 1).astype(np.float32)
    for w in (100, 101, 297):
        y = F.window_normalization(x, window_len=w).cpu().numpy()
        np.testing.assert_allclose(y, O.window_normalization(x, 1, w, True), rtol=2e-4, atol=2e-4)


def test_map_stage_with_normalisation(built_lib):
    from lidbox_b200.data import tf_utils
    rng = np.random.default_rng(21)
    sig = (rng.standard_normal((2, 32000)) * 0.1).astype(np.float32)
    rates = np.array([16000, 16000])
    kw = dict(feat_scale_kwargs={"min": 0.0, "max": 1.0, "axis": None},
              window_norm_kwargs={"window_len": 100, "normalize_variance": True})
    X = tf_utils.extract_features(sig, rates, "logmelspectrogram", {}, {}, {}, {}, kw["feat_scale_kwargs"],
                                  kw["window_norm_kwargs"]).cpu().numpy()
    ref = O.extract_features(sig, rates, "logmelspectrogram", **kw)
    np.testing.assert_allclose(X, ref, rtol=2e-3, atol=2e-3)


def test_vad_parity_and_reference_properties(audio):
    # SURVEY §8(f) row 2.  Masks must be identical to the oracle's wherever the frame RMS is not within 1e-5 (relative)
    # of the threshold: fp32 summation order is not comparable across implementations exactly at a tie.
    g = np.load(os.path.join(GOLDEN, "wav_fixtures.npz"))
    for pcm in g["pcm"]:
        s = pcm.astype(np.float32) / np.float32(32768.0)
        vad = audio.framewise_rms_energy_vad_decisions(s, 16000, 25)
        assert vad.dtype == torch.bool and bool(vad.all())                       # tests/test_features_audio.py:175-179
        assert audio.remove_silence(s, 16000).shape == s.shape                   # :183-188
    z = np.zeros(3 * 16000, np.float32)
    assert not bool(audio.framewise_rms_energy_vad_decisions(z, 16000, 25).any())    # :180-181
    assert audio.remove_silence(z, 16000).numel() == 0                           # :189-191
    pos, length = audio.run_length_encoding(np.array([1, 1, 1, 2, 2, 2, 3, 4, 5, 6, 6, 7]))
    assert pos.tolist() == [0, 3, 6, 7, 8, 9, 11] and length.tolist() == [3, 3, 1, 1, 1, 2, 1]   # :166-169 exact KAT
    rng = np.random.default_rng(30)
    x = rng.normal(0, 5, size=(7, 9))
    np.testing.assert_allclose(audio.root_mean_square(x, axis=-1).cpu().numpy(), O.root_mean_square(x), rtol=1e-5)
    # speech-like bursts separated by silence, batched
    B, N = 6, 48000
    sig = np.zeros((B, N), np.float32)
    for b in range(B):
        for _ in range(4):
            a, n = int(rng.integers(0, N - 8000)), int(rng.integers(800, 8000))
            tone = np.sin(2 * np.pi * rng.uniform(100, 3000) * np.arange(n) / 16000.0)
            sig[b, a:a + n] += (rng.uniform(0.05, 0.8) * tone).astype(np.float32)
        sig[b] += 1e-4 * rng.standard_normal(N).astype(np.float32)
    for (ms, nonspeech, strength) in ((10, 0, 0.05), (10, 300, 0.1), (25, 100, 0.5), (30, 0, 1.0)):
        dec = audio.batched_rms_vad(sig, 16000, ms, min_non_speech_ms=nonspeech, strength=strength).cpu().numpy()
        for b in range(B):
            ref, margin = O.framewise_rms_energy_vad_decisions(sig[b], 16000, ms, nonspeech, strength, return_margin=True)
            assert dec[b].shape == ref.shape
            if (margin > 1e-5).all():
                assert (dec[b] == ref).all()
    out, lengths = audio.batched_remove_silence(sig, 16000)
    for b in range(B):
        ref = O.remove_silence(sig[b], 16000)
        assert int(lengths[b]) == ref.size
        assert np.array_equal(out[b, :ref.size].cpu().numpy(), ref)              # compaction is a pure copy: bit-exact


def test_pcm16_input_is_bit_identical(audio):
    # 16-bit PCM decoded inside the kernel (x / 32768) must equal the float32 path on the decoded signal exactly
    g = np.load(os.path.join(GOLDEN, "wav_fixtures.npz"))
    pcm = torch.from_numpy(g["pcm"])
    a = audio.logmelspectrograms(pcm, 16000)
    b = audio.logmelspectrograms(pcm.float() / 32768.0, 16000)
    assert torch.equal(a, b)
    odd = pcm[:, 3:7777].contiguous()          # unaligned rows
    assert torch.equal(audio.logmelspectrograms(odd, 16000), audio.logmelspectrograms(odd.float() / 32768.0, 16000))


def test_mfcc_map_stage(built_lib):
    from lidbox_b200.data import tf_utils
    from lidbox_b200.features import audio as a
    rng = np.random.default_rng(22)
    sig = (rng.standard_normal((3, 16000)) * 0.1).astype(np.float32)
    rates = np.array([16000] * 3)
    X = tf_utils.extract_features(sig, rates, "mfcc", {}, {}, {}, {}, {}, {}).cpu().numpy()
    ref = O.extract_features(sig, rates, "mfcc")
    assert X.shape == ref.shape == (3, 98, 12)
    np.testing.assert_allclose(X, ref, rtol=1e-4, atol=1e-4)
    X2 = tf_utils.extract_features(sig, rates, "mfcc", {}, {}, {"coef_begin": 0, "coef_end": 20}).cpu().numpy()
    np.testing.assert_allclose(X2, O.extract_features(sig, rates, "mfcc", mfcc_kwargs={"coef_begin": 0, "coef_end": 20}),
                               rtol=1e-4, atol=1e-4)
    lm = rng.standard_normal((2, 5, 40)).astype(np.float32)
    np.testing.assert_allclose(a.mfccs_from_log_mel_spectrograms(lm).cpu().numpy(),
                               O.mfccs_from_log_mel_spectrograms(lm), rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# round 2: persistent packed-fp32 kernel — staging paths, multi-pass CTAs, direct bf16 hand-off, host entry point
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N", [(3, 16000), (2, 16001), (5, 16002), (1, 16399), (700, 4000), (2, 400), (1, 559),
                                 (1, 560)])
def test_logmel_staging_paths_and_persistent_passes(audio, B, N):
    """N % 4 == 0 takes the bulk-copy (TMA) prefetch path, other lengths the plain-load fallback; B = 700 x 0.25 s
    gives every persistent CTA several runs (more items than 3 CTAs x 148 SMs); N = 400 / 559 / 560 are the
    one-frame and two-frame edges."""
    sig = _signals(B, N, seed=N)
    ref = O.logmel(sig, 16000, dtype=np.float64)
    out = audio.logmelspectrograms(sig, 16000).cpu().numpy()
    assert out.shape == ref.shape
    assert_logmel_close(out, ref)


@pytest.mark.parametrize("N", [16000, 16008, 16004, 16001])
def test_logmel_pcm16_staging_paths(audio, N):
    """16-bit PCM: N % 8 == 0 -> bulk copy of the raw samples + in-place expansion; otherwise plain loads."""
    pcm = np.clip(np.round(_signals(3, N, seed=7) * 32768.0), -32768, 32767).astype(np.int16)
    ref = O.logmel(pcm.astype(np.float32) / 32768.0, 16000, dtype=np.float64)
    out = audio.logmelspectrograms(torch.from_numpy(pcm), 16000).cpu().numpy()
    assert_logmel_close(out, ref)


def test_logmel_nonfinite_input_stays_in_its_utterance(audio):
    """Persistent CTAs reuse the staging buffer: a NaN in one utterance must not leak into another one's frames."""
    sig = _signals(40, 8000, seed=3)
    sig[17, 4000] = np.nan
    out = audio.logmelspectrograms(sig, 16000).cpu().numpy()
    bad = ~np.isfinite(out).all(axis=(1, 2))
    assert bad[17] and bad.sum() == 1


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_logmel_direct_handoff_equals_packing_pass(audio, precision):
    """logmelspectrograms(out=model.feature_sink(B, T)) writes bf16 rows straight into the first frame layer's
    buffer; the model output must be bit-identical to feeding the fp32 features through the packing pass."""
    from lidbox_b200.models import xvector
    sig = _signals(6, 16000, seed=11)
    m = xvector.create((98, 40), 5, precision=precision, seed=1)
    feats = audio.logmelspectrograms(sig, 16000)
    ref = m(feats).cpu().numpy()
    m._buffers(6, 98, False)["X"][0].zero_()
    sink = audio.logmelspectrograms(sig, 16000, out=m.feature_sink(6, 98))
    out = m(sink).cpu().numpy()
    np.testing.assert_array_equal(out, ref)
    pcm = torch.from_numpy(np.clip(np.round(sig * 32768.0), -32768, 32767).astype(np.int16))
    out16 = m(audio.logmelspectrograms(pcm, 16000, out=m.feature_sink(6, 98))).cpu().numpy()
    ref16 = m(audio.logmelspectrograms(pcm, 16000)).cpu().numpy()
    np.testing.assert_array_equal(out16, ref16)


def test_logmel_host_entry_point(audio, built_lib):
    """lbx_logmel_f32_host: pageable host signals in, host log-mel out (the binding a non-torch caller uses)."""
    import ctypes
    from lidbox_b200 import _lib
    sig = _signals(3, 16000, seed=5)
    T = 98
    out = np.empty((3, T, 40), np.float32)
    d_sig = torch.empty((3, 16000), dtype=torch.float32, device="cuda")
    d_out = torch.empty((3, T, 40), dtype=torch.float32, device="cuda")
    d_tab = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    rc = _lib.lib().lbx_logmel_f32_host(sig.ctypes.data_as(ctypes.c_void_p), 3, 16000, 16000, 25, 10, 512, 2.0, 40, 0.0,
                                        8000.0, 1, 1e-6, out.ctypes.data_as(ctypes.c_void_p), _lib.ptr(d_sig),
                                        _lib.ptr(d_out), _lib.ptr(d_tab), d_tab.numel(),
                                        _lib.stream_ptr(d_sig.device))
    _lib.check(rc)
    assert_logmel_close(out, O.logmel(sig, 16000, dtype=np.float64))
