"""Drop-in for lidbox/features/mel_ops.py: the (deliberately bug-compatible) mel weight matrix.

The table is a host-side constant: built once per (num_mel_bins, num_spectrogram_bins, sample_rate, fmin, fmax) by
lbx_mel_weight_matrix (fp32, same evaluation order as mel_ops.py:11-75, including `_linspace` dividing by `num`),
band-compressed, uploaded and cached per device.
"""
import ctypes
import threading

import numpy as np
import torch

from .. import _lib

_cache = {}
_cache_lock = threading.Lock()


def _weight_matrix_host(num_mel_bins, num_spectrogram_bins, sample_rate, lower_edge_hertz, upper_edge_hertz):
    W = np.empty((int(num_spectrogram_bins), int(num_mel_bins)), np.float32)
    _lib.check(_lib.lib().lbx_mel_weight_matrix(int(num_mel_bins), int(num_spectrogram_bins), int(sample_rate),
                                                float(lower_edge_hertz), float(upper_edge_hertz),
                                                W.ctypes.data_as(ctypes.c_void_p)))
    return W


def linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000, lower_edge_hertz=125.0,
                                upper_edge_hertz=3800.0, dtype=torch.float32, name=None):
    """lidbox/features/mel_ops.py:28-75 -> [num_spectrogram_bins, num_mel_bins] tensor (host, float32)."""
    W = _weight_matrix_host(num_mel_bins, num_spectrogram_bins, sample_rate, lower_edge_hertz, upper_edge_hertz)
    return torch.from_numpy(W).to(dtype)


class MelBands:
    """Band-compressed filterbank resident on one device: per mel bin the contiguous run of non-zero FFT bins."""

    def __init__(self, W, device):
        n_bins, n_mel = W.shape
        start = np.empty(n_mel, np.int32)
        length = np.empty(n_mel, np.int32)
        off = np.empty(n_mel, np.int32)
        packed = np.empty(n_bins * n_mel, np.float32)
        as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        n = _lib.lib().lbx_mel_pack_bands(as_p(W), n_bins, n_mel, as_p(start), as_p(length), as_p(off), as_p(packed))
        if n < 0:
            _lib.check(n)
        self.n_bins, self.n_mel, self.n_packed = n_bins, n_mel, int(n)
        self.start = torch.from_numpy(start).to(device)
        self.len = torch.from_numpy(length).to(device)
        self.off = torch.from_numpy(off).to(device)
        self.w = torch.from_numpy(packed[:max(n, 1)].copy()).to(device)


def mel_bands(num_mel_bins, num_spectrogram_bins, sample_rate, fmin, fmax, device):
    key = (int(num_mel_bins), int(num_spectrogram_bins), int(sample_rate), float(fmin), float(fmax), str(device))
    with _cache_lock:
        bands = _cache.get(key)
        if bands is None:
            W = _weight_matrix_host(*key[:5])
            bands = _cache[key] = MelBands(W, device)
    return bands
