"""GPU tests (-m gpu) of the two frame-layer extensions north_star names and lidbox/models/xvector.py does not have
(SURVEY.md §0.1): dilated causal Conv1D (Keras semantics) and BatchNormalization behind the activation
(lidbox/models/xvector_2d.py:41-43 order, inference).  The reference here is torch's own conv1d / batch_norm in fp64 on
the bf16-rounded operands the kernels see — independent code, not the oracle restating itself."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TDNN = [(64, 5, 1, 1), (64, 3, 1, 2), (72, 3, 1, 3), (64, 1, 1, 1), (96, 1, 1, 1)]      # (filters, k, stride, dilation)


def _model(xv, layers, T, n_out, **kw):
    frames = [xv.frame_layer(f, k, s, dilation_rate=d, name="frame%d" % (i + 1), batch_norm=kw.pop("bn", False) and i < 2)
              for i, (f, k, s, d) in enumerate(layers)]
    segments = [xv.segment_layer(64, name="segment1"), xv.segment_layer(64, name="segment2")]
    return xv.XVector((T, 40), n_out, frames=frames, segments=segments, **kw)


def _torch_forward(w, layers, x, bn=None, bf16_acts=True):
    """fp64 forward with bf16 rounding of every stored activation (what the bf16 path keeps in HBM)."""
    rnd = (lambda t: t.to(torch.bfloat16).to(torch.float64)) if bf16_acts else (lambda t: t)
    h = rnd(x.double()).transpose(1, 2)                                       # [B, C, T]
    for i, (f, k, s, d) in enumerate(layers):
        name = "frame%d" % (i + 1)
        W = rnd(w[name + "/kernel"].double()).permute(2, 1, 0)                # Keras [k, Cin, Cout] -> torch [Cout, Cin, k]
        h = F.conv1d(F.pad(h, ((k - 1) * d, 0)), W, w[name + "/bias"].double(), stride=s, dilation=d)
        h = torch.relu(h)
        if bn and name + "_bn/gamma" in bn:
            g, b, mu, var = (bn[name + "_bn/" + key].double() for key in ("gamma", "beta", "moving_mean", "moving_variance"))
            h = F.batch_norm(h, mu, var, g, b, training=False, eps=1e-3)
        h = rnd(h)
    mean = h.mean(2)
    std = torch.sqrt(torch.clamp(((h - mean[:, :, None]) ** 2).mean(2), min=1e-10))
    z = rnd(torch.cat([mean, std], 1))
    for name in ("segment1", "segment2"):
        z = rnd(torch.relu(z @ rnd(w[name + "/kernel"].double()) + w[name + "/bias"].double()))
    return z @ rnd(w["outputs/kernel"].double()) + w["outputs/bias"].double()


def _weights(m, seed):
    g = torch.Generator().manual_seed(seed)
    w = {k: torch.as_tensor(v) for k, v in m.get_weights().items()}
    for k in w:
        if k.endswith("/bias"):
            w[k] = torch.randn(w[k].shape, generator=g) * 0.05
        elif k.endswith("_bn/gamma") or k.endswith("_bn/moving_variance"):
            w[k] = torch.rand(w[k].shape, generator=g) + 0.5
        elif "_bn/" in k:
            w[k] = torch.randn(w[k].shape, generator=g) * 0.3
    m.set_weights({k: v.numpy() for k, v in w.items()})
    return w


@pytest.mark.parametrize("T", [37, 120])
def test_dilated_tdnn_forward_matches_torch_conv1d(built_lib, T):
    from lidbox_b200.models import xvector as xv
    m = _model(xv, TDNN, T, 6, precision="bf16", seed=1, head="none")
    w = _weights(m, 2)
    x = torch.randn(5, T, 40, generator=torch.Generator().manual_seed(3))
    got = m(x.numpy()).double().cpu()
    ref = _torch_forward(w, TDNN, x)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max()) + 1e-3      # bf16 activations: rounding flips only


def test_keras_rejects_stride_with_dilation(built_lib):
    from lidbox_b200.models import xvector as xv
    with pytest.raises(ValueError):
        xv.frame_layer(64, 3, 2, dilation_rate=2)


def test_batch_norm_inference_after_relu(built_lib):
    from lidbox_b200.models import xvector as xv
    layers = [(64, 5, 1, 1), (64, 3, 2, 1), (64, 3, 1, 2), (96, 1, 1, 1)]
    m = _model(xv, layers, 50, 4, precision="bf16", seed=4, head="none", bn=True)
    w = _weights(m, 5)
    x = torch.randn(3, 50, 40, generator=torch.Generator().manual_seed(6))
    got = m(x.numpy()).double().cpu()
    ref = _torch_forward(w, layers, x, bn=w)
    assert float((got - ref).abs().max()) < 2e-2 * float(ref.abs().max()) + 1e-3
    # and the affine really is applied: without it the result differs visibly
    ref0 = _torch_forward(w, layers, x)
    assert float((ref0 - ref).abs().max()) > 20 * float((got - ref).abs().max())
    with pytest.raises(NotImplementedError):
        m.loss_and_grads(x.numpy(), np.zeros(3, dtype=np.int64))


def test_dilated_tdnn_gradients_match_torch_autograd(built_lib):
    """loss_and_grads of the dilated TDNN vs torch autograd (fp64, same bf16 rounding points as the forward pass):
    per-tensor cosine >= 0.995 and max-norm error <= 5 %, the bar of the undilated gradient tests."""
    from lidbox_b200.models import xvector as xv
    T, B, n_out = 61, 12, 4
    m = _model(xv, TDNN, T, n_out, precision="bf16", seed=7)
    w = _weights(m, 8)
    x = torch.randn(B, T, 40, generator=torch.Generator().manual_seed(9))
    y = torch.arange(B) % n_out
    loss = m.loss_and_grads(x.numpy(), y.numpy()).cpu().double()
    got = {}
    for ly in m.layers:
        gw = m._w_view(ly, m.grads)[:, :ly["N"]].cpu().double()
        if ly["kind"] == "frame":
            gw = gw.view(ly["k"], ly["c_in"], ly["N"])[:, :ly["c_in_real"]]
        got[ly["name"] + "/kernel"] = gw
        got[ly["name"] + "/bias"] = m.grads[ly["b_off"]:ly["b_off"] + ly["N"]].cpu().double()
    wd = {k: v.double().requires_grad_(True) for k, v in w.items()}

    class _Ste(torch.autograd.Function):              # bf16 rounding with a straight-through gradient
        @staticmethod
        def forward(ctx, t):
            return t.to(torch.bfloat16).to(torch.float64)

        @staticmethod
        def backward(ctx, g):
            return g

    h = _Ste.apply(x.double()).transpose(1, 2)
    for i, (f, k, s, d) in enumerate(TDNN):
        name = "frame%d" % (i + 1)
        h = F.conv1d(F.pad(h, ((k - 1) * d, 0)), _Ste.apply(wd[name + "/kernel"]).permute(2, 1, 0), wd[name + "/bias"],
                     stride=s, dilation=d)
        h = _Ste.apply(torch.relu(h))
    mean = h.mean(2)
    std = torch.sqrt(torch.clamp(((h - mean[:, :, None]) ** 2).mean(2), min=1e-10))
    z = _Ste.apply(torch.cat([mean, std], 1))
    for name in ("segment1", "segment2"):
        z = _Ste.apply(torch.relu(z @ _Ste.apply(wd[name + "/kernel"]) + wd[name + "/bias"]))
    logits = z @ _Ste.apply(wd["outputs/kernel"]) + wd["outputs/bias"]
    per = -torch.log_softmax(logits, 1)[torch.arange(B), y]
    per.mean().backward()
    assert float((loss - per.detach()).abs().max()) < 2e-2
    for k in got:
        a, b = got[k].flatten(), wd[k].grad.flatten()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.995, (k, cos)
        assert float((a - b).abs().max()) < 0.05 * float(b.abs().max()) + 1e-7, k
