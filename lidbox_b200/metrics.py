"""Drop-in for lidbox/metrics.py: average detection cost C_avg (Li, Ma & Lee 2013, eq. 32) with device-side counters.

    cavg = AverageDetectionCost(N, thresholds, C_miss=1.0, C_fa=1.0, P_tar=0.5)    # metrics.py:19-47
    cavg.update_state(true_positives_onehot, predictions)                           # metrics.py:52-72
    cavg.result()                                                                   # metrics.py:74-99 -> min over thresholds
    SparseAverageDetectionCost(...).update_state(true_sparse, predictions)          # metrics.py:104-109

The counters tp/fn [N,Th] and fp_pairs/tn_pairs [N,N,Th] live in HBM and are updated by one kernel per batch
(lbx_cavg_update_f32); result() is one kernel and returns a device scalar, so an evaluation loop never synchronises.
"""
import numpy as np
import torch

from . import _lib


def _f32(x, device):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(device, torch.float32).contiguous()


class AverageDetectionCost:
    def __init__(self, N, thresholds, C_miss=1.0, C_fa=1.0, P_tar=0.5, name="C_avg", device=None, **kwargs):
        if int(N) < 2:
            raise ValueError("C_avg is undefined for less than 2 classes.")                     # metrics.py:21
        thr = np.asarray(thresholds, dtype=np.float32)
        if thr.ndim != 1 or thr.size < 1:
            raise ValueError("Thresholds must be an array of decision scores.")                 # metrics.py:22
        self.name = name
        self.N, self.num_thresholds = int(N), int(thr.size)
        self.device = _lib.require_cuda(device)
        self.thresholds = torch.as_tensor(thr).to(self.device)
        self.C_miss, self.C_fa, self.P_tar = float(C_miss), float(C_fa), float(P_tar)
        n, t = self.N, self.num_thresholds
        # one allocation, four views: reset_states is a single memset
        self._state = torch.zeros(2 * n * t + 2 * n * n * t, dtype=torch.float32, device=self.device)
        self.tp = self._state[:n * t].view(n, t)
        self.fn = self._state[n * t:2 * n * t].view(n, t)
        self.fp_pairs = self._state[2 * n * t:2 * n * t + n * n * t].view(n, n, t)
        self.tn_pairs = self._state[2 * n * t + n * n * t:].view(n, n, t)

    def reset_states(self):
        self._state.zero_()

    reset_state = reset_states

    def _update(self, onehot, labels, predictions):
        pred = _f32(predictions, self.device)
        if pred.dim() != 2 or pred.shape[1] != self.N:
            raise ValueError("predictions must have shape [batch, %d], got %s" % (self.N, tuple(pred.shape)))
        B = pred.shape[0]
        if (onehot is not None and tuple(onehot.shape) != (B, self.N)) or (labels is not None and labels.numel() != B):
            raise ValueError("labels and predictions disagree on the batch size / number of classes")
        _lib.check(_lib.lib().lbx_cavg_update_f32(
            _lib.ptr(onehot) if onehot is not None else None, _lib.ptr(labels) if labels is not None else None,
            _lib.ptr(pred), B, self.N, _lib.ptr(self.thresholds), self.num_thresholds, _lib.ptr(self.tp),
            _lib.ptr(self.fn), _lib.ptr(self.fp_pairs), _lib.ptr(self.tn_pairs), _lib.stream_ptr(self.device)))

    def update_state(self, true_positives, predictions, **kwargs):
        """Dense (float one-hot) labels [B, N] and scores [B, N]."""
        self._update(_f32(true_positives, self.device), None, predictions)

    def result_per_threshold(self):
        out = torch.empty(self.num_thresholds + 1, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().lbx_cavg_result_f32(_lib.ptr(self.tp), _lib.ptr(self.fn), _lib.ptr(self.fp_pairs),
                                                  _lib.ptr(self.tn_pairs), self.N, self.num_thresholds, self.C_miss,
                                                  self.C_fa, self.P_tar, _lib.ptr(out),
                                                  out.data_ptr() + 4 * self.num_thresholds,
                                                  _lib.stream_ptr(self.device)))
        return out

    def result(self):
        """Smallest C_avg over the given thresholds (0-dim device tensor)."""
        return self.result_per_threshold()[self.num_thresholds]

    def __call__(self, true_positives, predictions):
        self.update_state(true_positives, predictions)
        return self.result()


class SparseAverageDetectionCost(AverageDetectionCost):
    def update_state(self, true_positives, predictions, **kwargs):
        """Sparse integer labels [B] (or [B, 1]) and scores [B, N]."""
        y = true_positives if isinstance(true_positives, torch.Tensor) else torch.as_tensor(np.asarray(true_positives))
        self._update(None, y.to(self.device, torch.int32).reshape(-1).contiguous(), predictions)
