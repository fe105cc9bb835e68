"""Drop-in for lidbox/losses.py: SparseAngularProximity on a hand-written CUDA kernel (forward + backward).

    loss_fn = SparseAngularProximity(N, D, delta_weight=1.0)
    loss = loss_fn(y_true_sparse, y_pred)        # scalar: mean over the batch (Keras SUM_OVER_BATCH_SIZE)
    per_sample = loss_fn.call(y_true_sparse, y_pred)   # [B]            (losses.py:25-40)
    theta = loss_fn.theta(z)                     # [B, N]               (losses.py:42-49)
    scores = loss_fn.predict(z)                  # -theta               (losses.py:51-52)

Labels of shape [B] or [B, 1] are accepted and read as one class index per sample (the semantics of the in-tree
self-test, losses.py:70,97; see SURVEY.md §8 A13 for the [B,1] broadcasting quirk that is NOT reproduced).
`call` is differentiable w.r.t. y_pred through torch.autograd (the backward is the same CUDA kernel).
"""
import numpy as np
import torch

from . import _lib


def _ap_launch(z, y, N, w, want_theta, want_loss, gloss):
    B, D = z.shape
    lib, st = _lib.lib(), _lib.stream_ptr(z.device)
    theta = torch.empty((B, N), dtype=torch.float32, device=z.device) if want_theta else None
    loss = torch.empty((B,), dtype=torch.float32, device=z.device) if want_loss else None
    grad = torch.empty((B, D), dtype=torch.float32, device=z.device) if gloss is not None else None
    _lib.check(lib.lbx_ap_loss(_lib.ptr(z), _lib.ptr(y), B, D, N, float(w), 0, None, _lib.ptr(theta), _lib.ptr(loss),
                               _lib.ptr(grad), None, D, _lib.ptr(gloss), 1.0, None, st))
    return theta, loss, grad


class _APFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, y, N, w):
        ctx.save_for_backward(z, y)
        ctx.N, ctx.w = N, w
        return _ap_launch(z, y, N, w, False, True, None)[1]

    @staticmethod
    def backward(ctx, gloss):
        z, y = ctx.saved_tensors
        grad = _ap_launch(z, y, ctx.N, ctx.w, False, False, gloss.to(torch.float32).contiguous())[2]
        return grad, None, None, None


class SparseAngularProximity:
    def __init__(self, N, D, delta_weight=1.0, name="AP", **kwargs):
        # losses.py:14-16 (tf.debugging asserts -> ValueError here)
        if N < 1:
            raise ValueError("Must have at least 1 class")
        if D < N:
            raise ValueError("Language vector dimension cannot be less than number of classes")
        if not delta_weight > 0:
            raise ValueError("Non-positive delta weight would cause correct classifications to have larger loss "
                             "values than incorrect classifications.")
        self.N, self.D, self.delta_weight, self.name = int(N), int(D), float(delta_weight), name

    def _z(self, y_pred):
        z = y_pred if isinstance(y_pred, torch.Tensor) else torch.as_tensor(np.asarray(y_pred))
        if z.dim() != 2 or z.shape[1] != self.D:
            raise ValueError("y_pred must have shape [batch_size, %d]" % self.D)
        return z.to(_lib.require_cuda(), torch.float32).contiguous()

    def _y(self, y_true_sparse, B, device):
        y = torch.as_tensor(np.asarray(y_true_sparse) if not isinstance(y_true_sparse, torch.Tensor) else y_true_sparse)
        y = y.to(device, torch.int32).reshape(-1).contiguous()
        if y.numel() != B:
            raise ValueError("expected one label per sample")
        return y

    def call(self, y_true_sparse, y_pred):
        z = self._z(y_pred)
        return _APFunction.apply(z, self._y(y_true_sparse, z.shape[0], z.device), self.N, self.delta_weight)

    def __call__(self, y_true_sparse, y_pred):
        return self.call(y_true_sparse, y_pred).mean()

    def theta(self, z):
        z = self._z(z)
        y = torch.zeros(z.shape[0], dtype=torch.int32, device=z.device)
        return _ap_launch(z, y, self.N, self.delta_weight, True, False, None)[0]

    def predict(self, z):
        return -self.theta(z)
