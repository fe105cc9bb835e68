"""Drop-in for the hot-path functions of lidbox/features/audio.py.

Same names, argument order, defaults and shape contracts as the reference; inputs may be NumPy arrays, Python
scalars or torch tensors (as the reference's tests pass them, tests/test_features_audio.py:121,139-142,151-152);
results are float32 torch tensors on the CUDA device.  Everything numeric is one C-ABI call into hand-written
sm_100a kernels (include/lidbox_b200.h).
"""
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
    kernel: [B, N] -> [B, T, num_mel_bins].  Not a reference name: it is what the map stage calls."""
    sig = _as_device_f32(signals, 2, "signals")
    L = ms_to_frames(sample_rate, frame_length_ms)
    step = ms_to_frames(sample_rate, frame_step_ms)
    B, N = sig.shape
    lib = _lib.lib()
    T = int(lib.lbx_num_frames(N, L, step)) if L >= 1 and step >= 1 else 0
    K = int(fft_length) // 2 + 1
    bands = mel_ops.mel_bands(num_mel_bins, K, sample_rate, fmin, fmax, sig.device)
    if out is None:
        out = torch.empty((B, T, int(num_mel_bins)), dtype=torch.float32, device=sig.device)
    ws_bytes = int(lib.lbx_logmel_workspace_bytes(B, N, L, step, int(fft_length), bands.n_mel))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=sig.device) if ws_bytes else None
    _lib.check(lib.lbx_logmel_f32(_lib.ptr(sig), B, N, L, step, int(fft_length), float(power), bands.n_mel,
                                  _lib.ptr(bands.start), _lib.ptr(bands.len), _lib.ptr(bands.off), _lib.ptr(bands.w),
                                  bands.n_packed, 1 if log else 0, float(eps), _lib.ptr(out), _lib.ptr(ws), ws_bytes,
                                  _lib.stream_ptr(sig.device)))
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
