"""Cross-checks of the CPU oracle against INDEPENDENT third-party code available in this image.

The oracle (oracle/lidbox_oracle.py) restates TensorFlow-2.3 semantics that the reference relies on but that are not
in /root/reference.  TensorFlow is not installable here, so these tests pin each restated semantic
against an implementation that was written by someone else:

  semantic (reference call site)                                  third party
  --------------------------------------------------------------  ------------------------------------------------
  periodic Hann window (audio.py:229 via tf.signal.stft)          scipy.signal.get_window(fftbins=True),
                                                                  torch.hann_window(periodic=True)
  framing T = 1 + (N-L)//step, no centring, window multiply,      scipy.signal.stft(boundary=None, padded=False),
  zero padding AT THE END to fft_length, rFFT (audio.py:229)      scipy.signal.ShortTimeFFT, transformers.audio_utils
  HTK mel scale + triangular filters in mel space                 transformers.audio_utils.mel_filter_bank(
  (mel_ops.py:23-25, 57-75)                                       mel_scale="htk", triangularize_in_mel_space=True)
  MFCC = DCT-II * rsqrt(2N) (tf_utils.py:183)                     scipy.fft.dct(type=2, norm=None / "ortho")
  power_to_db (audio.py:167-174)                                  transformers.audio_utils.power_to_db
  causal strided Conv1D (xvector.py:38-39, Keras)                 torch.nn.functional.conv1d over an independently
                                                                  written left pad; torch.nn.Unfold-free loop form
  Dense / log-softmax / population variance (xvector.py:25-65)    torch.nn.functional.linear / log_softmax / torch.var
  AP-loss gradient (losses.py:25-49)                              torch autograd of an independently written
                                                                  forward + finite differences

What stays "TF from memory" after these tests is listed in DESIGN.md §1(c).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import lidbox_oracle as O  # noqa: E402


def _signals(B, N, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(N) / 16000.0
    f = rng.uniform(100.0, 4000.0, size=(B, 1))
    return (0.5 * np.sin(2 * np.pi * f * t) + 0.05 * rng.standard_normal((B, N))).astype(np.float32)


# ------------------------------------------------------------------------------------------------ window
@pytest.mark.parametrize("L", [1, 2, 3, 160, 399, 400, 401, 512, 1102])
def test_hann_periodic_vs_scipy_and_torch(L):
    """TF's raised-cosine windows use n = L + periodic*even - 1 (the in-tree copy of that formula is blackman_window,
    lidbox/features/audio.py:206-211, which cites window_ops.py of TF v2.3.1): even L -> the DFT-periodic window
    (denominator L), odd L -> the SYMMETRIC window (denominator L-1) even though periodic=True."""
    import torch
    from scipy.signal import get_window
    w = O.hann_window(L, np.float64)
    if L == 1:
        assert w.tolist() == [1.0]
    elif L % 2 == 0:
        np.testing.assert_allclose(w, get_window("hann", L, fftbins=True), atol=1e-15)
        np.testing.assert_allclose(w, torch.hann_window(L, periodic=True, dtype=torch.float64).numpy(), atol=1e-15)
    else:
        np.testing.assert_allclose(w, get_window("hann", L, fftbins=False), atol=1e-15)
        np.testing.assert_allclose(w, torch.hann_window(L, periodic=False, dtype=torch.float64).numpy(), atol=1e-15)


def test_hann_is_not_the_symmetric_window():
    from scipy.signal import get_window
    assert np.abs(O.hann_window(400, np.float64) - get_window("hann", 400, fftbins=False)).max() > 1e-3


# ------------------------------------------------------------------------------------------------ STFT
@pytest.mark.parametrize("L,step,nfft,N", [(400, 160, 512, 16000), (400, 160, 512, 16399), (400, 160, 512, 400),
                                           (256, 128, 256, 8000), (300, 100, 1024, 5000), (401, 77, 512, 3000)])
def test_stft_vs_scipy_stft(L, step, nfft, N):
    """scipy.signal.stft with boundary=None, padded=False frames exactly like tf.signal.frame(pad_end=False), applies the
    window to the L-sample segment and lets rfft zero-pad at the END to nfft.  Its only difference is the 1/sum(w)
    'spectrum' scaling, undone here."""
    from scipy.signal import get_window, stft
    x = _signals(3, N, seed=L + N).astype(np.float64)
    w = get_window("hann", L, fftbins=(L % 2 == 0))            # odd L: TF's "periodic" window is the symmetric one
    _, _, Z = stft(x, window=w, nperseg=L, noverlap=L - step, nfft=nfft, boundary=None, padded=False,
                   return_onesided=True, scaling="spectrum")
    Z = np.moveaxis(Z, -1, 1) * w.sum()                         # [B, T, K]
    S = O.stft(x, L, step, nfft, np.float64)
    assert S.shape == Z.shape == (3, 1 + (N - L) // step, nfft // 2 + 1)
    np.testing.assert_allclose(S, Z, atol=1e-9 * np.abs(Z).max())


def test_stft_vs_scipy_shorttimefft():
    """A second, structurally different scipy implementation: ShortTimeFFT slices by hop index p; frames fully inside
    the signal are p in [p0, p1) with the window anchored at its centre sample m_num_mid."""
    from scipy.signal import ShortTimeFFT, get_window
    L, step, nfft, N = 400, 160, 512, 8000
    x = _signals(1, N, seed=5)[0].astype(np.float64)
    w = get_window("hann", L, fftbins=True)
    sft = ShortTimeFFT(w, hop=step, fs=16000, mfft=nfft, fft_mode="onesided", scale_to=None, phase_shift=None)
    Z = sft.stft(x)                                             # [K, P]; slice p covers samples p*hop - m_num_mid ...
    S = O.stft(x[None], L, step, nfft, np.float64)[0]           # [T, K]
    T = S.shape[0]
    for t in (0, 1, T // 2, T - 1):
        p = (t * step + sft.m_num_mid) / step                   # the slice whose window starts at sample t*step
        if abs(p - round(p)) > 1e-12:
            continue
        zp = Z[:, int(round(p)) - sft.p_min]
        np.testing.assert_allclose(np.abs(S[t]), np.abs(zp), atol=1e-9 * np.abs(zp).max())


def test_spectrogram_vs_transformers_audio_utils():
    """transformers.audio_utils.spectrogram(center=False) is a third framing + window + rfft implementation."""
    au = pytest.importorskip("transformers.audio_utils")
    x = _signals(2, 16000, seed=11)
    w = au.window_function(400, "hann", periodic=True)
    for b in range(2):
        P = au.spectrogram(x[b].astype(np.float64), w, frame_length=400, hop_length=160, fft_length=512, power=2.0,
                           center=False, dtype=np.float64)   # [K, T]
        S = O.spectrograms(x[b:b + 1], 16000, dtype=np.float64)[0]
        assert S.shape == (98, 257)
        np.testing.assert_allclose(S, P.T, rtol=2e-6, atol=1e-9 * P.max())   # transformers keeps a complex64 buffer


def test_torch_stft_centres_the_window_and_is_not_the_reference_semantic():
    """torch.stft pads the 400-sample window on BOTH sides inside the 512-sample frame, so it is not an
    oracle for tf.signal.stft — recorded here so nobody 'fixes' the oracle with it."""
    import torch
    x = torch.from_numpy(_signals(1, 4000, seed=2))
    Zt = torch.stft(x, n_fft=512, hop_length=160, win_length=400, window=torch.hann_window(400, periodic=True),
                    center=False, return_complex=True)          # [1, K, T']
    S = O.stft(x.numpy(), 400, 160, 512)
    assert Zt.shape[-1] != S.shape[1] or not np.allclose(np.abs(Zt[0].T.numpy()), np.abs(S[0]), rtol=1e-3)


# ------------------------------------------------------------------------------------------------ mel
def test_mel_matrix_formula_family_vs_transformers(monkeypatch):
    """With the reference's two off-by-one divisions replaced by a true linspace, the restatement must be the
    canonical tf.signal.linear_to_mel_weight_matrix, which transformers reproduces with mel_scale="htk",
    triangularize_in_mel_space=True.  This pins _hertz_to_mel, the triangle slopes, the clamp and the zeroed DC row;
    the off-by-one divisors themselves are literal in lidbox/features/mel_ops.py:16,40-41,49-55."""
    au = pytest.importorskip("transformers.audio_utils")

    def true_linspace(start, stop, num, dtype=np.float64):
        return np.linspace(dtype(start), dtype(stop), num, dtype=dtype)

    monkeypatch.setattr(O, "_linspace", true_linspace)
    for (M, K, sr, lo, hi) in [(40, 257, 16000, 0.0, 8000.0), (20, 129, 8000, 125.0, 3800.0), (64, 513, 16000, 20.0, 7600.0)]:
        W = O.linear_to_mel_weight_matrix(M, K, sr, lo, hi, np.float64)
        Wt = au.mel_filter_bank(K, M, lo, hi, sr, norm=None, mel_scale="htk", triangularize_in_mel_space=True)
        assert W.shape == Wt.shape == (K, M)
        np.testing.assert_allclose(W, Wt, atol=1e-9)


def test_mel_matrix_reference_bug_is_kept():
    """The bug-compatible matrix differs from the canonical one: top edge below Nyquist, bins 241..256 empty."""
    au = pytest.importorskip("transformers.audio_utils")
    W = O.linear_to_mel_weight_matrix(40, 257, 16000, 0.0, 8000.0)
    Wt = au.mel_filter_bank(257, 40, 0.0, 8000.0, 16000, norm=None, mel_scale="htk", triangularize_in_mel_space=True)
    assert np.abs(W - Wt).max() > 1e-2
    assert (W[241:] == 0).all() and (W != 0).sum() == 464


def test_hertz_to_mel_vs_transformers():
    au = pytest.importorskip("transformers.audio_utils")
    f = np.linspace(0.0, 8000.0, 101)
    np.testing.assert_allclose(O._hertz_to_mel(f, np.float64), au.hertz_to_mel(f, mel_scale="htk"), rtol=2e-4)
    # transformers uses 2595*log10(1+f/700); 1127*ln(1+f/700) differs from it by 1127.0 vs 1127.01048 (HTK rounding)
    np.testing.assert_allclose(O._hertz_to_mel(f, np.float64), 1127.0 * np.log1p(f / 700.0), rtol=1e-12)


# ------------------------------------------------------------------------------------------------ MFCC / dB
def test_mfcc_vs_scipy_dct():
    from scipy.fft import dct
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 7, 40))
    ours = O.mfccs_from_log_mel_spectrograms(x)
    np.testing.assert_allclose(ours, dct(x, type=2, norm=None, axis=-1) / np.sqrt(2.0 * 40), atol=1e-12)
    ortho = dct(x, type=2, norm="ortho", axis=-1)               # ortho differs only in coefficient 0 (x 1/sqrt(2))
    np.testing.assert_allclose(ours[..., 1:], ortho[..., 1:], atol=1e-12)
    np.testing.assert_allclose(ours[..., 0], ortho[..., 0] * np.sqrt(2.0), atol=1e-12)


def test_power_to_db_vs_transformers():
    """audio.py:167-174 is librosa's power_to_db(ref=np.max) with a 20x (not 10x) factor; transformers implements the
    10x librosa formula independently."""
    au = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(4)
    S = (rng.random((2, 30, 17)) ** 4).astype(np.float32)
    ref = au.power_to_db(S.astype(np.float64), reference=float(S.max()), min_value=1e-10, db_range=40.0)
    np.testing.assert_allclose(O.power_to_db(S, amin=1e-10, top_db=80.0), 2.0 * ref, atol=2e-4)


# ------------------------------------------------------------------------------------------------ TDNN pieces
def _indep_causal_conv(x, kernel, bias, stride):
    """Loop statement of SURVEY App. A.8 written without looking at the oracle: out[b,t,o] = bias[o] +
    sum_{j<k} sum_c xpad[b, t*s + j, c] W[j,c,o] with k-1 zero frames on the left, T_out = ceil(T / s)."""
    B, T, C = x.shape
    k, _, Co = kernel.shape
    xpad = np.concatenate([np.zeros((B, k - 1, C)), x], axis=1)
    T_out = -(-T // stride)
    out = np.zeros((B, T_out, Co))
    for t in range(T_out):
        for j in range(k):
            out[:, t] += xpad[:, t * stride + j] @ kernel[j]
    return out + bias


@pytest.mark.parametrize("k,s,T", [(5, 1, 17), (3, 2, 17), (3, 2, 18), (3, 3, 19), (1, 1, 5), (3, 4, 21), (3, 4, 1),
                                   (5, 1, 1), (2, 3, 7)])
def test_causal_strided_conv_vs_torch_conv1d_and_loops(k, s, T):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(k * 100 + s * 10 + T)
    x = rng.standard_normal((2, T, 6))
    W = rng.standard_normal((k, 6, 4))
    b = rng.standard_normal(4)
    ours = O.conv1d_causal(x, W, b, s, relu=False)
    # torch: NCW layout, weight [C_out, C_in, k] (cross-correlation, like Keras), explicit left pad of k-1
    xt = F.pad(torch.from_numpy(x).permute(0, 2, 1), (k - 1, 0))
    yt = F.conv1d(xt, torch.from_numpy(W).permute(2, 1, 0), torch.from_numpy(b), stride=s).permute(0, 2, 1).numpy()
    assert ours.shape == yt.shape == (2, -(-T // s), 4)
    np.testing.assert_allclose(ours, yt, atol=1e-10)
    np.testing.assert_allclose(ours, _indep_causal_conv(x, W, b, s), atol=1e-10)
    np.testing.assert_allclose(O.conv1d_causal(x, W, b, s, relu=True), np.maximum(yt, 0.0), atol=1e-10)


def test_xvector_forward_vs_torch_modules():
    """The whole forward (frame1..5, pooling, segments, outputs, log_softmax) rebuilt from torch.nn.functional
    primitives with Keras-layout weights — independent of the oracle's NumPy code path."""
    import torch
    import torch.nn.functional as F
    params = O.xvector_init(23, 7, seed=3, dtype=np.float64, bias_scale=0.1)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((3, 50, 23))
    ours = O.xvector_forward(params, x)
    emb = O.xvector_forward(params, x, embedding=True)
    h = torch.from_numpy(x).permute(0, 2, 1)
    for name, k, s in [("frame1", 5, 1), ("frame2", 3, 2), ("frame3", 3, 3), ("frame4", 1, 1), ("frame5", 1, 1)]:
        W = torch.from_numpy(params[name + "/kernel"]).permute(2, 1, 0)
        h = F.relu(F.conv1d(F.pad(h, (k - 1, 0)), W, torch.from_numpy(params[name + "/bias"]), stride=s))
    mean = h.mean(dim=2)
    std = torch.sqrt(torch.clamp(h.var(dim=2, unbiased=False), min=1e-10))
    p = torch.cat([mean, std], dim=1)
    e = F.linear(p, torch.from_numpy(params["segment1/kernel"]).T, torch.from_numpy(params["segment1/bias"]))
    np.testing.assert_allclose(emb, e.numpy(), atol=1e-9)
    h2 = F.relu(F.linear(F.relu(e), torch.from_numpy(params["segment2/kernel"]).T,
                         torch.from_numpy(params["segment2/bias"])))
    out = F.log_softmax(F.linear(h2, torch.from_numpy(params["outputs/kernel"]).T,
                                 torch.from_numpy(params["outputs/bias"])), dim=1)
    np.testing.assert_allclose(ours, out.numpy(), atol=1e-9)


def test_stats_pooling_vs_numpy_population_statistics():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 33, 10))
    x[0, :, 3] = 2.5                                            # zero variance -> sqrt(clip) = 1e-5
    out = O.stats_pooling(x)
    np.testing.assert_allclose(out[:, :10], x.mean(1), atol=1e-12)
    np.testing.assert_allclose(out[:, 10:], np.sqrt(np.clip(x.var(1, ddof=0), 1e-10, None)), atol=1e-12)
    assert abs(out[0, 13] - 1e-5) < 1e-12


def test_ap_loss_gradient_vs_autograd_and_finite_differences():
    """losses.py:25-49 written again from the paper's formula (per-sample sum over l' != y of sigmoid(w (theta_y -
    theta_l'))) with torch ops; its autograd gradient and a central finite difference must agree with the oracle."""
    import torch
    rng = np.random.default_rng(6)
    B, D, N, w = 5, 12, 7, 1.5
    z = rng.standard_normal((B, D))
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    y = rng.integers(0, N, size=B)
    zt = torch.tensor(z, requires_grad=True)
    theta = torch.acos(zt[:, :N])
    ty = theta.gather(1, torch.tensor(y)[:, None])
    terms = torch.sigmoid(w * (ty - theta))
    mask = torch.ones(B, N)
    mask[torch.arange(B), torch.tensor(y)] = 0.0
    per_sample = (terms * mask).sum(1)
    np.testing.assert_allclose(O.ap_loss_per_sample(y, z, N, w), per_sample.detach().numpy(), atol=1e-12)
    per_sample.mean().backward()
    g = zt.grad.numpy()
    assert np.abs(g[:, N:]).max() == 0.0
    eps = 1e-6
    for (b, l) in [(0, 0), (1, 3), (4, 6), (2, int(y[2]))]:
        zp, zm = z.copy(), z.copy()
        zp[b, l] += eps
        zm[b, l] -= eps
        fd = (O.ap_loss(y, zp, N, w) - O.ap_loss(y, zm, N, w)) / (2 * eps)
        assert abs(fd - g[b, l]) < 1e-6 * max(1.0, abs(g[b, l]))
    lt = O.torch_ap_loss(torch.tensor(y), torch.tensor(z, requires_grad=False), N, w)
    assert abs(float(lt.mean()) - float(per_sample.mean())) < 1e-12


def test_adam_keras_formula_vs_torch_adam():
    """Keras Adam (eps = 1e-7 added to sqrt(v_hat) after bias correction folded into lr_t) vs torch.optim.Adam:
    lr_t = lr sqrt(1-b2^t)/(1-b1^t);  p -= lr_t m / (sqrt(v) + eps)  — torch uses eps' = eps / sqrt(1-b2^t) in the
    same algebra, so with eps -> 0 they coincide; the test pins the moment updates and the bias correction."""
    import torch
    rng = np.random.default_rng(2)
    p0 = rng.standard_normal(50)
    grads = [rng.standard_normal(50) for _ in range(4)]
    pt = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([pt], lr=1e-3, betas=(0.9, 0.999), eps=1e-30)
    p, m, v = p0.copy(), np.zeros(50), np.zeros(50)
    for t, g in enumerate(grads, 1):
        pt.grad = torch.tensor(g)
        opt.step()
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        p = p - lr_t * m / (np.sqrt(v) + 1e-30)
    np.testing.assert_allclose(p, pt.detach().numpy(), atol=1e-12)
