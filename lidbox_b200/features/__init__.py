"""Drop-in for lidbox/features/__init__.py: feature_scaling, cmn, cmvn, window_normalization on CUDA kernels
(csrc/normalize.cu).  Same names, argument order and defaults as the reference; NumPy arrays or torch tensors in,
float32 torch CUDA tensors out."""
import numpy as np
import torch

from .. import _lib


def _prep(X):
    t = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X))
    if t.is_complex() or t.dtype == torch.bool:
        raise TypeError("expected a real-valued tensor")
    return t.to(_lib.require_cuda(), torch.float32).contiguous()


def _view(shape, axis):
    axis = axis if axis >= 0 else axis + len(shape)
    if not 0 <= axis < len(shape):
        raise ValueError("axis %d out of range for rank %d" % (axis, len(shape)))
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def _normalize(X, axis, mode, lo=0.0, hi=1.0):
    X = _prep(X)
    out = torch.empty_like(X)
    outer, R, inner = _view(X.shape, axis)
    _lib.check(_lib.lib().lbx_normalize_axis_f32(_lib.ptr(X), _lib.ptr(out), outer, R, inner, mode, float(lo), float(hi),
                                                 _lib.stream_ptr(X.device)))
    return out


def feature_scaling(X, min, max, axis=None):
    """lidbox/features/__init__.py:5-9 — scale X over `axis` (None: the whole tensor) into [min, max]."""
    if axis is None:
        X = _prep(X)
        out = torch.empty_like(X)
        ws = torch.empty(2, dtype=torch.int32, device=X.device)
        _lib.check(_lib.lib().lbx_feature_scaling_all_f32(_lib.ptr(X), _lib.ptr(out), X.numel(), float(min), float(max),
                                                          _lib.ptr(ws), _lib.stream_ptr(X.device)))
        return out
    return _normalize(X, axis, 2, min, max)


def cmn(X, axis=1):
    """lidbox/features/__init__.py:12-20 — centre the means over `axis` (rank-3 input, like the reference signature)."""
    X = _prep(X)
    if X.dim() != 3:
        raise ValueError("cmn expects a rank-3 tensor [B, T, F]")
    return _normalize(X, axis, 0)


def cmvn(X, axis=1):
    """lidbox/features/__init__.py:22-32 — zero mean, unit (population) variance over `axis`; x/0 -> 0."""
    X = _prep(X)
    if X.dim() != 3:
        raise ValueError("cmvn expects a rank-3 tensor [B, T, F]")
    return _normalize(X, axis, 1)


def window_normalization(X, axis=1, window_len=-1, normalize_variance=True):
    """lidbox/features/__init__.py:34-67 — sliding-window mean (and variance) normalisation over the time axis."""
    X = _prep(X)
    if X.dim() != 3:
        raise ValueError("window_normalization expects a rank-3 tensor [B, T, F]")
    B, T, F = X.shape
    if window_len == -1 or T <= window_len:
        return cmvn(X, axis=axis) if normalize_variance else cmn(X, axis=axis)
    if axis != 1:
        raise NotImplementedError("sliding windows are implemented over the time axis (axis=1), as the reference tests")
    out = torch.empty_like(X)
    _lib.check(_lib.lib().lbx_window_normalization_f32(_lib.ptr(X), _lib.ptr(out), B, T, F, int(window_len),
                                                       int(bool(normalize_variance)), _lib.stream_ptr(X.device)))
    return out
