"""Drop-in for the hot-path functions of lidbox/features/audio.py.

Same names, argument order, defaults and shape contracts as the reference; inputs may be NumPy arrays, Python
scalars or torch tensors (as the reference's tests pass them, tests/test_features_audio.py:121,139-142,151-152);
results are float32 torch tensors on the CUDA device.  Everything numeric is one C-ABI call into hand-written
sm_100a kernels (include/lidbox_b200.h).
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from . import mel_ops


def _as_device_f32(x, rank, name):
    dev = _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dim() != rank:
        raise ValueError("%s must have rank %d, got shape %s" % (name, rank, tuple(t.shape)))
    if t.is_complex() or t.dtype == torch.bool:
        raise TypeError("%s must be real-valued, got %s" % (name, t.dtype))
    if not t.is_cuda:
        t = t.to(dev, non_blocking=True)
    return t.to(torch.float32).contiguous()


def ms_to_frames(sample_rate, ms):
    """lidbox/features/audio.py:185-189."""
    return int(_lib.lib().lbx_ms_to_frames(int(sample_rate), int(ms)))


def fft_frequencies(sample_rate, n_fft):
    """lidbox/features/audio.py:150-159 (tiny host-side table; returned on the host)."""
    return torch.linspace(0.0, float(int(sample_rate) // 2), 1 + int(n_fft) // 2, dtype=torch.float64).to(torch.float32)


def log10(x):
    """lidbox/features/audio.py:162-164."""
    x = torch.as_tensor(x)
    return torch.log(x) / torch.log(torch.tensor(10.0, dtype=x.dtype, device=x.device))


def spectrograms(signals, sample_rate, frame_length_ms=25, frame_step_ms=10, power=2.0, fft_length=512):
    """lidbox/features/audio.py:219-230: [B, N] -> [B, T, fft_length/2 + 1]."""
    sig = _as_device_f32(signals, 2, "signals")
    L = ms_to_frames(sample_rate, frame_length_ms)
    step = ms_to_frames(sample_rate, frame_step_ms)
    B, N = sig.shape
    lib = _lib.lib()
    T = int(lib.lbx_num_frames(N, L, step)) if L >= 1 and step >= 1 else 0
    out = torch.empty((B, T, int(fft_length) // 2 + 1), dtype=torch.float32, device=sig.device)
    _lib.check(lib.lbx_spectrogram_f32(_lib.ptr(sig), B, N, L, step, int(fft_length), float(power), _lib.ptr(out),
                                       _lib.stream_ptr(sig.device)))
    return out


def linear_to_mel(spectrograms, sample_rate, num_mel_bins=40, fmin=0.0, fmax=8000.0, _log_mode=0, _eps=1e-6):
    """lidbox/features/audio.py:247-261: [B, T, K] -> [B, T, num_mel_bins]."""
    S = _as_device_f32(spectrograms, 3, "spectrograms")
    B, T, K = S.shape
    bands = mel_ops.mel_bands(num_mel_bins, K, sample_rate, fmin, fmax, S.device)
    out = torch.empty((B, T, int(num_mel_bins)), dtype=torch.float32, device=S.device)
    _lib.check(_lib.lib().lbx_linear_to_mel_f32(_lib.ptr(S), B * T, K, bands.n_mel, _lib.ptr(bands.start),
                                                _lib.ptr(bands.len), _lib.ptr(bands.off), _lib.ptr(bands.w),
                                                bands.n_packed, int(_log_mode), float(_eps), _lib.ptr(out),
                                                _lib.stream_ptr(S.device)))
    return out


def logmelspectrograms(signals, sample_rate, frame_length_ms=25, frame_step_ms=10, power=2.0, fft_length=512,
                       num_mel_bins=40, fmin=0.0, fmax=8000.0, log=True, eps=1e-6, out=None):
    """Fused spectrograms -> linear_to_mel -> ln(x + 1e-6) (the chain of lidbox/data/tf_utils.py:172-178) in one
    kernel: [B, N] -> [B, T, num_mel_bins].  Not a reference name: it is what the map stage calls.
    An int16 torch tensor is taken as 16-bit PCM and decoded on the fly (x / 32768, as read_wav does).
    `out` may be a float32 tensor [B, T, num_mel_bins], or the feature sink of a model
    (`XVector.feature_sink(B, T)`): the rows are then written as bf16 straight into the zero-left-padded activation
    buffer of the first frame layer (no fp32 round trip, no packing pass) and the sink is returned."""
    dev = _lib.require_cuda()
    if isinstance(signals, torch.Tensor) and signals.dtype == torch.int16:
        if signals.dim() != 2:
            raise ValueError("signals must have rank 2, got shape %s" % (tuple(signals.shape),))
        sig, sig_dtype = signals.to(dev, non_blocking=True).contiguous(), _lib.I16
    else:
        sig, sig_dtype = _as_device_f32(signals, 2, "signals"), _lib.F32
    L = ms_to_frames(sample_rate, frame_length_ms)
    step = ms_to_frames(sample_rate, frame_step_ms)
    B, N = sig.shape
    lib = _lib.lib()
    T = int(lib.lbx_num_frames(N, L, step)) if L >= 1 and step >= 1 else 0
    K = int(fft_length) // 2 + 1
    bands = mel_ops.mel_bands(num_mel_bins, K, sample_rate, fmin, fmax, sig.device)
    d = _lib.LogmelDesc()
    d.sig, d.sig_dtype, d.B, d.N = sig.data_ptr(), sig_dtype, B, N
    d.frame_length, d.frame_step, d.fft_length, d.power = L, step, int(fft_length), float(power)
    d.n_mel, d.n_packed = bands.n_mel, bands.n_packed
    d.band_start, d.band_len, d.band_off, d.band_w = (bands.start.data_ptr(), bands.len.data_ptr(),
                                                      bands.off.data_ptr(), bands.w.data_ptr())
    d.log_mode, d.eps = 1 if log else 0, float(eps)
    ws = None
    if hasattr(out, "sink_spec"):
        hi, lo, utt_pitch, row_pitch = out.sink_spec(B, T, bands.n_mel)
        d.out, d.out_lo, d.out_dtype, d.out_utt_pitch, d.out_row_pitch = hi, lo, _lib.BF16, utt_pitch, row_pitch
    else:
        if out is None:
            out = torch.empty((B, T, int(num_mel_bins)), dtype=torch.float32, device=sig.device)
        elif tuple(out.shape) != (B, T, int(num_mel_bins)) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % ((B, T, int(num_mel_bins)),))
        d.out, d.out_dtype = out.data_ptr(), _lib.F32
        ws_bytes = int(lib.lbx_logmel_workspace_bytes(B, N, L, step, int(fft_length), bands.n_mel))
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=sig.device)
            d.workspace, d.workspace_bytes = ws.data_ptr(), ws_bytes
    _lib.check(lib.lbx_logmel_ex(ctypes.byref(d), _lib.stream_ptr(sig.device)))
    return out


def mfccs_from_log_mel_spectrograms(log_mel, coef_begin=0, coef_end=None):
    """tf.signal.mfccs_from_log_mel_spectrograms(X)[..., coef_begin:coef_end] as used by the map stage
    (lidbox/data/tf_utils.py:180-185): [B, T, M] -> [B, T, coef_end - coef_begin]."""
    X = _as_device_f32(log_mel, 3, "log_mel")
    B, T, M = X.shape
    coef_end = M if coef_end is None else min(int(coef_end), M)
    coef_begin = min(int(coef_begin), coef_end)
    out = torch.empty((B, T, coef_end - coef_begin), dtype=torch.float32, device=X.device)
    _lib.check(_lib.lib().lbx_mfcc_f32(_lib.ptr(X), B * T, M, coef_begin, coef_end, _lib.ptr(out),
                                       _lib.stream_ptr(X.device)))
    return out


def db_to_power(S):
    """lidbox/features/audio.py:177-181: pow(10, S / 20) on a rank-3 float32 tensor."""
    S = _as_device_f32(S, 3, "S")
    out = torch.empty_like(S)
    _lib.check(_lib.lib().lbx_db_to_power_f32(_lib.ptr(S), S.numel(), _lib.ptr(out), _lib.stream_ptr(S.device)))
    return out


def power_to_db(S, amin=1e-10, top_db=80.0):
    """lidbox/features/audio.py:167-174 (max over the whole tensor, batch included)."""
    S = _as_device_f32(S, 3, "S")
    out = torch.empty_like(S)
    ws = torch.empty(4, dtype=torch.float32, device=S.device)
    _lib.check(_lib.lib().lbx_power_to_db_f32(_lib.ptr(S), S.numel(), float(amin), float(top_db), _lib.ptr(out),
                                              _lib.ptr(ws), _lib.stream_ptr(S.device)))
    return out


def assert_all_finite(X, message):
    """tf.debugging.assert_all_finite as used after every stage of tf_utils.extract_features (:173-194)."""
    flag = torch.zeros(1, dtype=torch.int32, device=X.device)
    Xc = X.contiguous()
    _lib.check(_lib.lib().lbx_check_finite_f32(_lib.ptr(Xc), Xc.numel(), _lib.ptr(flag), _lib.stream_ptr(X.device)))
    if int(flag.item()) != 0:
        raise FloatingPointError(message)


# ------------------------------------------------------------------------------------------------------------
# Energy VAD (lidbox/features/audio.py:264-353) — SURVEY.md §8(f) row 2.  The kernels are batched over utterances
# ([B, N]); the reference's single-utterance signatures are served with B = 1.
# ------------------------------------------------------------------------------------------------------------
def root_mean_square(x, axis=-1):
    """lidbox/features/audio.py:262-271 (rank-2 input, RMS over `axis`)."""
    t = _as_device_f32(x, 2, "x")
    if axis in (0, -2):
        t = t.t().contiguous()
    elif axis not in (1, -1):
        raise ValueError("axis out of range for a rank-2 tensor")
    rows, n = t.shape
    out = torch.empty((rows,), dtype=torch.float32, device=t.device)
    if n == 0:
        return out.fill_(float("nan"))
    _lib.check(_lib.lib().lbx_row_rms_f32(_lib.ptr(t), rows, n, _lib.ptr(out), _lib.stream_ptr(t.device)))
    return out


def run_length_encoding(v):
    """lidbox/features/audio.py:273-283: (start positions, lengths) of the runs of equal values of an int vector.
    Index bookkeeping on a handful of elements: done on the host."""
    v = np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v).reshape(-1)
    if v.size == 0:
        return torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)
    change = np.flatnonzero(v[1:] != v[:-1])
    i = np.concatenate(([-1], change, [v.size - 1]))
    pos = np.concatenate(([0], np.cumsum(i[1:] - i[:-1])))
    return torch.from_numpy(pos[:-1].astype(np.int64)), torch.from_numpy((pos[1:] - pos[:-1]).astype(np.int64))


def invert_too_short_consecutive_false(mask, min_length):
    """lidbox/features/audio.py:285-296: runs of False shorter than min_length become True."""
    m = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask).astype(bool).reshape(-1)
    if min_length == 0 or m.size == 0:
        return torch.from_numpy(m.copy())
    pos, lengths = run_length_encoding(m.astype(np.int32))
    keep = np.logical_or(m[pos.numpy()], lengths.numpy() < min_length)
    return torch.from_numpy(np.repeat(keep, lengths.numpy()))


def batched_rms_vad(signals, sample_rate, frame_step_ms, min_non_speech_ms=0, strength=0.05, min_rms_threshold=1e-3):
    """framewise_rms_energy_vad_decisions for a batch [B, N] in two kernel launches -> bool [B, N // frame_step]."""
    sig = _as_device_f32(signals, 2, "signals")
    B, N = sig.shape
    step = ms_to_frames(sample_rate, frame_step_ms)
    if step < 1:
        raise ValueError("frame step must be at least one sample")
    F = N // step
    # audio.py:325: cast(ms_to_frames(sr, min_non_speech_ms) / frame_step, int64) — true division, then truncation
    min_frames = int(ms_to_frames(sample_rate, min_non_speech_ms) / step)
    dec = torch.zeros((B, F), dtype=torch.uint8, device=sig.device)
    ws = torch.empty((B, max(F, 1)), dtype=torch.float32, device=sig.device)
    _lib.check(_lib.lib().lbx_rms_vad_f32(_lib.ptr(sig), B, N, step, float(strength), float(min_rms_threshold),
                                          min_frames, _lib.ptr(dec), _lib.ptr(ws), _lib.stream_ptr(sig.device)))
    return dec.bool()


def framewise_rms_energy_vad_decisions(signal, sample_rate, frame_step_ms, min_non_speech_ms=0, strength=0.05,
                                       min_rms_threshold=1e-3, time_axis=0):
    """lidbox/features/audio.py:298-329 (rank-1 signal; True = voiced)."""
    if time_axis != 0:
        raise NotImplementedError("only time_axis=0 (the rank-1 signature of the reference) is implemented")
    s = torch.as_tensor(np.asarray(signal) if not isinstance(signal, torch.Tensor) else signal)
    if s.dim() != 1:
        raise ValueError("signal must have rank 1")
    return batched_rms_vad(s[None], sample_rate, frame_step_ms, min_non_speech_ms, strength, min_rms_threshold)[0]


def batched_remove_silence(signals, rate, window_ms=10, min_non_speech_ms=300, vad=None):
    """remove_silence / apply_vad for a batch: returns (out [B, N] with the voiced windows compacted to the front,
    lengths [B] in samples).  `vad` may carry precomputed decisions [B, F] (steps.py:191-198)."""
    sig = _as_device_f32(signals, 2, "signals")
    B, N = sig.shape
    window = (int(window_ms) * int(rate)) // 1000               # audio.py:341 (integer arithmetic, not ms_to_frames)
    if vad is None:
        vad = batched_rms_vad(sig, rate, window_ms, min_non_speech_ms=min_non_speech_ms, strength=0.1)
    dec = vad.to(sig.device, torch.uint8).contiguous()
    F = dec.shape[1]
    if window < 1 or F * window > N:
        raise ValueError("VAD decisions do not match the signal length")
    out = torch.zeros_like(sig)
    lengths = torch.zeros((B,), dtype=torch.int64, device=sig.device)
    offs = torch.empty((B, max(F, 1)), dtype=torch.int64, device=sig.device)
    _lib.check(_lib.lib().lbx_vad_compact_f32(_lib.ptr(sig), B, N, window, _lib.ptr(dec), F, _lib.ptr(out),
                                              _lib.ptr(lengths), _lib.ptr(offs), _lib.stream_ptr(sig.device)))
    return out, lengths


def remove_silence(signal, rate, window_ms=10, min_non_speech_ms=300):
    """lidbox/features/audio.py:331-353 (rank-1 signal -> voiced samples)."""
    s = torch.as_tensor(np.asarray(signal) if not isinstance(signal, torch.Tensor) else signal)
    if s.dim() != 1:
        raise ValueError("signal must have rank 1")
    out, lengths = batched_remove_silence(s[None], rate, window_ms, min_non_speech_ms)
    return out[0, :int(lengths[0].item())]
