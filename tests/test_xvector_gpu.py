"""GPU parity tests (-m gpu) of the x-vector TDNN, statistics pooling, cross-entropy and angular-proximity loss
against the fp64 CPU oracle (oracle/lidbox_oracle.py), all through the C-ABI.

Tolerances
  precision="fp32" (bf16x3 tensor-core accumulation, BASELINE config 2): normwise 1e-4 — the north-star bound
  precision="bf16" (training configs 3-4): checked twice —
      (a) against the oracle restated with bfloat16 rounding at the same storage points (emulate_bf16=True):
          loss 1e-3, every gradient tensor cosine >= 0.999 (measured 0.99935 .. 1.0) and max-norm 0.2 (accumulation order, bf16 rounding ties);
      (b) against the plain fp64 oracle: loss 3e-2, gradient cosine >= 0.98 per tensor (the distance between bf16
          mixed-precision training and fp64 through 8 layers; measured 0.990 .. 0.999998)
  pooling / losses in fp32: 1e-5
"""
import os

import numpy as np
import pytest
import torch

from oracle import lidbox_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def xv(built_lib):
    assert torch.cuda.is_available()
    from lidbox_b200.models import xvector
    return xvector


def _nw(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _oracle_grads(params, x, y, loss="xent", N=None, w=1.0, emulate_bf16=False):
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    xt = torch.tensor(x, dtype=torch.float64)
    if loss == "xent":
        lp = O.torch_xvector_forward(tp, xt, emulate_bf16=emulate_bf16)
        l = -lp[torch.arange(len(y)), torch.tensor(y)].mean()
    else:
        z = O.torch_xvector_forward(tp, xt, l2_normalize=True, emulate_bf16=emulate_bf16)
        l = O.torch_ap_loss(torch.tensor(y), z, N, w)
    l.backward()
    return float(l.detach()), {k: v.grad.numpy() for k, v in tp.items()}


GRAD_NW_BOUND = 0.2     # normwise bound of the gradient checks vs the bf16-emulating oracle


def _check_grads(m, g_ref, min_cos, max_nw):
    grads = m.grads.cpu().numpy()
    for ly in m.layers:
        gw = grads[ly["w_off"]:ly["w_off"] + ly["K"] * ly["ldw"]].reshape(ly["K"], ly["ldw"])
        assert not gw[:, ly["N"]:].any()                 # pitch padding never receives a gradient
        gw = gw[:, :ly["N"]]
        gb = grads[ly["b_off"]:ly["b_off"] + ly["N"]]
        rw = g_ref[ly["name"] + "/kernel"].reshape(ly["K"], ly["N"])
        rb = g_ref[ly["name"] + "/bias"]
        for got, ref, what in ((gw, rw, "kernel"), (gb, rb, "bias")):
            cos = (got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30)
            assert cos > min_cos, "%s/%s cosine %.6f" % (ly["name"], what, cos)
            if max_nw is not None:
                assert _nw(got, ref) < max_nw, "%s/%s normwise %.4f" % (ly["name"], what, _nw(got, ref))


def test_golden_small_fp32(xv):
    g = np.load(os.path.join(GOLDEN, "xvector_small.npz"))
    m = xv.create((None, 24), 5)
    m.set_weights(O.xvector_init(24, 5, seed=11, bias_scale=0.05))
    logp = m(g["x"]).cpu().numpy()
    assert logp.shape == (3, 5)
    np.testing.assert_allclose(logp, g["logp"], rtol=1e-4, atol=1e-4)
    emb = xv.as_embedding_extractor(m)(g["x"]).cpu().numpy()
    assert _nw(emb, g["emb"]) < 1e-4


def test_config2_embedding_fp32(xv):
    # BASELINE config 2: batch 64 x 2 s (198 frames x 40 mel), forward to the 512-d segment1 embedding, fp32
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 198, 40)).astype(np.float32) * 3.0 - 5.0
    params = O.xvector_init(40, 4, seed=0, bias_scale=0.02)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    ref, acts = O.xvector_forward(p64, x.astype(np.float64), embedding=True, return_activations=True)
    m = xv.create((198, 40), 4)
    m.set_weights(params)
    assert m.count_params() == 4508124 + 513 * 4
    logp = m(x).cpu().numpy()
    np.testing.assert_allclose(logp, O.xvector_forward(p64, x.astype(np.float64)), rtol=1e-4, atol=1e-4)
    emb = xv.as_embedding_extractor(m)(x).cpu().numpy()
    assert emb.shape == (64, 512)
    assert _nw(emb, ref) < 1e-4


def test_reference_edge_shapes(xv):
    # /root/reference/tests/test_models.py:104-107 + lidbox/testutil.py:29-35: B,T,F >= 1, outputs 1..100
    rng = np.random.default_rng(1)
    for (B, T, F, n_out) in ((1, 1, 1, 1), (2, 7, 3, 100), (10, 400, 100, 4), (3, 2, 40, 7), (5, 13, 9, 2)):
        x = rng.uniform(-1e3, 1e3, size=(B, T, F)).astype(np.float32)
        m = xv.create(x.shape[1:], n_out)
        for training in (False, True):
            y = m(x, training=training).cpu().numpy()
            assert y.shape == (B, n_out) and not np.isnan(y).any()
        ref = O.xvector_forward({k: v.astype(np.float64) for k, v in m.get_weights().items()}, x.astype(np.float64))
        np.testing.assert_allclose(y, ref, rtol=1e-4, atol=2e-4 * max(1.0, np.abs(ref).max()))


def test_stats_pooling_layer(xv):
    rng = np.random.default_rng(2)
    pool = xv.GlobalMeanStddevPooling1D()
    for (B, T, C) in ((3, 33, 1500), (2, 1, 7), (4, 400, 64)):
        x = rng.standard_normal((B, T, C)).astype(np.float32) * 2 + 1
        out = pool(x).cpu().numpy()
        np.testing.assert_allclose(out, O.stats_pooling(x.astype(np.float64)), rtol=1e-5, atol=1e-5)
    # constant input: variance 0 -> std = sqrt(1e-10) (xvector.py:22,34)
    out = pool(np.full((1, 5, 3), 2.5, np.float32)).cpu().numpy()
    np.testing.assert_allclose(out[0, 3:], 1e-5, rtol=1e-5)


def test_channel_dropout_training_only(xv):
    x = np.random.default_rng(3).standard_normal((4, 50, 40)).astype(np.float32)
    m = xv.create((50, 40), 6, channel_dropout_rate=0.5)
    a, b = m(x, training=False), m(x, training=False)
    assert torch.allclose(a, b, rtol=0, atol=1e-5)      # split-K fp32 atomics: summation order is not fixed
    c = m(x, training=True)
    assert not torch.allclose(a, c) and torch.isfinite(c).all()


def _dropout_mask(m, x):
    """Reads the packed bf16 input of the first frame layer after a training-mode forward and returns the per
    (sample, channel) keep mask, asserting SpatialDropout1D semantics (xvector.py:50-51): a dropped channel is zero at
    EVERY frame of the sample, a kept one equals x / (1 - rate) (rounded to bf16)."""
    B, T, F = x.shape
    bufs = m._buffers(B, T, False)
    geo = bufs["geo"]
    X0 = bufs["X"][0][:B * geo.Tpad[0]].float().view(B, geo.Tpad[0], m.Fp)[:, geo.pad[0]:geo.pad[0] + T, :F].cpu().numpy()
    scale = 1.0 / (1.0 - m.channel_dropout_rate)
    want = torch.tensor(x * scale).to(torch.bfloat16).float().numpy()
    dropped = (X0 == 0).all(axis=1)                              # [B, F]
    kept = np.isclose(X0, want, rtol=1e-2, atol=1e-6).all(axis=1)
    assert (dropped | kept).all(), "a channel is neither fully dropped nor fully kept and rescaled"
    assert not (dropped & kept).any()
    return ~dropped


def test_spatial_dropout_semantics_and_rate(xv):
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((16, 30, 40)) + 3.0).astype(np.float32)      # no exact zeros in the input
    m = xv.create((30, 40), 6, channel_dropout_rate=0.5, seed=5)
    m(x, training=True)
    k1 = _dropout_mask(m, x)
    n = k1.size                                                         # 640 Bernoulli(0.5) draws: +- 5 sigma
    assert abs((~k1).sum() - 0.5 * n) < 5 * np.sqrt(n * 0.25)
    m(x, training=True)
    k2 = _dropout_mask(m, x)
    assert (k1 != k2).mean() > 0.3                                      # a fresh mask on every call
    m(x, training=False)
    assert _dropout_mask.__name__ == "_dropout_mask"
    k3 = (m._buffers(16, 30, False)["X"][0][:, :40] == 0).all().item()
    assert not k3                                                       # inference: nothing is dropped


def test_dropout_mask_changes_between_cuda_graph_replays(xv):
    """The mask index is a device-side counter advanced by a kernel, so replays of ONE captured graph draw different
    masks (a host-side counter would be frozen into the graph)."""
    rng = np.random.default_rng(9)
    B, T = 8, 30
    x = torch.tensor((rng.standard_normal((B, T, 40)) + 3.0).astype(np.float32), device="cuda")
    y = torch.tensor(np.arange(B) % 4, dtype=torch.int32, device="cuda")
    m = xv.create((T, 40), 4, channel_dropout_rate=0.5, precision="bf16", seed=2)
    m.configure_optimizer(lr=1e-4)
    step = xv.GraphedTrainStep(m, x, y)

    def mask():
        bufs = m._buffers(B, T, True)
        geo = bufs["geo"]
        X0 = bufs["X"][0][:B * geo.Tpad[0]].float().view(B, geo.Tpad[0], m.Fp)[:, geo.pad[0]:geo.pad[0] + T, :40]
        return (X0 == 0).all(dim=1).cpu().numpy()
    step()
    torch.cuda.synchronize()
    a = mask()
    step()
    torch.cuda.synchronize()
    b = mask()
    assert 0.2 < a.mean() < 0.8 and (a != b).mean() > 0.3


def test_graphed_step_copies_and_loss_read_back(xv):
    """GraphedTrainStep(copies=..., loss_host=...): the host-to-device transfer of a later batch and the read-back of the
    per-sample losses are nodes of the replayed graph — after a replay + synchronise the pinned loss buffer holds this
    replay's losses and the staged batch has arrived in its device buffer; the training result equals the plain step."""
    rng = np.random.default_rng(12)
    B, T = 8, 40
    x = torch.tensor(rng.standard_normal((B, T, 40)).astype(np.float32), device="cuda")
    y = torch.tensor(np.arange(B) % 4, dtype=torch.int32, device="cuda")
    staged_host = torch.tensor(rng.standard_normal((B, 1000)).astype(np.float32)).pin_memory()
    staged_dev = torch.zeros((B, 1000), device="cuda")
    loss_host = torch.full((B,), float("nan")).pin_memory()
    m = xv.create((T, 40), 4, precision="bf16", seed=2)
    m.configure_optimizer(lr=1e-3)
    ref = xv.create((T, 40), 4, precision="bf16", seed=2)
    ref.configure_optimizer(lr=1e-3)
    step = xv.GraphedTrainStep(m, x, y, copies=[(staged_dev, staged_host)], loss_host=loss_host, warmup=0)
    staged_dev.zero_()
    out = step()
    torch.cuda.synchronize()
    expect = ref.train_step(x, y)
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss_host.numpy(), expect.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out.cpu().numpy(), loss_host.numpy(), rtol=0, atol=0)
    assert torch.equal(staged_dev.cpu(), staged_host)
    n = ref.params.numel()
    assert float((m.params[:n] - ref.params).abs().max()) < 1e-5          # same update up to fp32 atomics order


@pytest.mark.parametrize("B,T,n_out", [(6, 37, 5), (32, 198, 4)])
def test_training_gradients_bf16_xent(xv, B, T, n_out):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, n_out, B)
    params = O.xvector_init(40, n_out, seed=3, bias_scale=0.05)
    m = xv.create((T, 40), n_out, precision="bf16")
    m.set_weights(params)
    per = m.loss_and_grads(x, y).cpu().numpy()
    loss_emu, g_emu = _oracle_grads(params, x, y, emulate_bf16=True)
    assert abs(per.mean() - loss_emu) < 1e-3 * max(1.0, abs(loss_emu))
    _check_grads(m, g_emu, 0.999, GRAD_NW_BOUND)
    loss_ref, g_ref = _oracle_grads(params, x, y)
    assert abs(per.mean() - loss_ref) < 3e-2 * max(1.0, abs(loss_ref))
    _check_grads(m, g_ref, 0.98, None)


def test_fused_xent_head_matches_unfused(xv, monkeypatch):
    # lbx_dense_xent_head (opt-in): output layer + log-softmax + cross-entropy + the layer's backward in one launch
    rng = np.random.default_rng(14)
    B, T, n_out = 19, 37, 4
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, n_out, B)
    params = O.xvector_init(40, n_out, seed=3, bias_scale=0.05)
    results = []
    for flag in ("0", "1"):
        monkeypatch.setenv("LBX_FUSED_HEAD", flag)
        m = xv.create((T, 40), n_out, precision="bf16")
        m.set_weights(params)
        per = m.loss_and_grads(x, y).cpu().numpy().copy()
        results.append((per, m.grads.cpu().numpy().copy(), m))
    (l0, g0, _), (l1, g1, m1) = results
    np.testing.assert_allclose(l1, l0, rtol=1e-5, atol=1e-6)
    cos = (g0 * g1).sum() / (np.linalg.norm(g0) * np.linalg.norm(g1))
    assert cos > 0.9999 and np.abs(g0 - g1).max() < 2e-2 * np.abs(g0).max()
    loss_emu, g_emu = _oracle_grads(params, x, y, emulate_bf16=True)
    assert abs(l1.mean() - loss_emu) < 1e-3 * max(1.0, abs(loss_emu))
    _check_grads(m1, g_emu, 0.999, GRAD_NW_BOUND)


def test_training_gradients_bf16_ap(xv):
    # BASELINE config 4 wiring at small size: x-vector -> 64-d L2-normalised vector -> SparseAngularProximity(50, 64)
    rng = np.random.default_rng(5)
    B, T, D, N = 8, 61, 64, 50
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, N, B)
    params = O.xvector_init(40, D, seed=4, bias_scale=0.05)
    m = xv.create((T, 40), D, precision="bf16", head="l2_normalize")
    m.set_weights(params)
    per = m.loss_and_grads(x, y, loss="ap", ap_classes=N).cpu().numpy()
    loss_emu, g_emu = _oracle_grads(params, x, y, loss="ap", N=N, emulate_bf16=True)
    assert abs(per.mean() - loss_emu) < 1e-3 * abs(loss_emu)
    _check_grads(m, g_emu, 0.999, GRAD_NW_BOUND)
    loss_ref, g_ref = _oracle_grads(params, x, y, loss="ap", N=N)
    assert abs(per.mean() - loss_ref) < 2e-2 * abs(loss_ref)
    _check_grads(m, g_ref, 0.98, None)


def test_training_reduces_loss(xv):
    # 4 synthetic classes separable by mean level; a few Adam steps must reduce the cross-entropy
    rng = np.random.default_rng(6)
    B, T = 64, 98
    y = np.arange(B) % 4
    x = (rng.standard_normal((B, T, 40)) + y[:, None, None] * 1.5).astype(np.float32)
    m = xv.create((T, 40), 4, precision="bf16", seed=1)
    m.configure_optimizer(lr=1e-3)
    first = float(m.train_step(x, y).mean())
    for _ in range(40):
        last = float(m.train_step(x, y).mean())
    assert np.isfinite(last) and last < 0.7 * first, (first, last)


def test_ap_loss_module(built_lib):
    from lidbox_b200.losses import SparseAngularProximity
    g = np.load(os.path.join(GOLDEN, "ap_loss.npz"))
    N, w = int(g["N"]), float(g["delta_weight"])
    lf = SparseAngularProximity(N, g["z"].shape[1], delta_weight=w)
    z = torch.tensor(g["z"], dtype=torch.float32, device="cuda", requires_grad=True)
    per = lf.call(g["y"], z)
    np.testing.assert_allclose(per.detach().cpu().numpy(), g["per_sample"], rtol=1e-5, atol=1e-5)
    loss = lf(g["y"].reshape(-1, 1), z)                     # [B,1] labels are squeezed
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5
    loss.backward()
    np.testing.assert_allclose(z.grad.cpu().numpy(), g["grad"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(lf.theta(g["z"]).cpu().numpy(), np.arccos(g["z"][:, :N]), atol=1e-5)
    np.testing.assert_allclose(lf.predict(g["z"]).cpu().numpy(), -np.arccos(g["z"][:, :N]), atol=1e-5)
    for bad in ((0, 4, 1.0), (5, 4, 1.0), (3, 4, 0.0)):      # losses.py:14-16
        with pytest.raises(ValueError):
            SparseAngularProximity(*bad)


def test_map_stage_extract_features(built_lib):
    from lidbox_b200.data import tf_utils
    rng = np.random.default_rng(7)
    sig = (rng.standard_normal((3, 16000)) * 0.1).astype(np.float32)
    rates = np.array([16000, 16000, 16000])
    for feattype in ("spectrogram", "melspectrogram", "logmelspectrogram", "db_spectrogram"):
        X = tf_utils.extract_features(sig, rates, feattype, {}, {}, {}, {}, {}, {}).cpu().numpy()
        ref = O.extract_features(sig, rates, feattype)
        assert X.shape == ref.shape
        if feattype == "db_spectrogram":
            np.testing.assert_allclose(X, ref, atol=2e-3)
        else:
            assert _nw(X, ref) < 1e-4
    X = tf_utils.extract_features(sig, rates, "logmelspectrogram", {"frame_length_ms": 20, "frame_step_ms": 5},
                                  {"num_mel_bins": 64, "fmin": 20.0, "fmax": 7600.0})
    assert X.shape == (3, 1 + (16000 - 320) // 80, 64)
    # an unknown feature type falls through the reference's if/elif chain: the spectrogram comes back (tf_utils.py:172-188)
    X = tf_utils.extract_features(sig, rates, "no-such-type").cpu().numpy()
    assert _nw(X, O.spectrograms(sig, 16000, dtype=np.float64)) < 1e-4
    from lidbox_b200.features import audio
    S = audio.power_to_db(audio.spectrograms(sig, 16000))
    np.testing.assert_allclose(audio.db_to_power(S).cpu().numpy(), O.db_to_power(S.cpu().numpy()), rtol=1e-5)
    with pytest.raises(ValueError):
        tf_utils.extract_features(sig[0], rates, "spectrogram")                       # rank != 2 (tf_utils.py:168)
    with pytest.raises(ValueError):
        tf_utils.extract_features(sig, np.array([16000, 8000, 16000]), "spectrogram")  # tf_utils.py:169
    bad = sig.copy(); bad[1, 5000] = np.nan
    with pytest.raises(FloatingPointError):
        tf_utils.extract_features(bad, rates, "logmelspectrogram")                    # tf_utils.py:173-194


# ---------------------------------------------------------------------------------------------------------------
# round 2: gradient parity at the BASELINE config sizes.  Bounds: cosine per tensor vs the
# bf16-emulating oracle, and the normwise error max|g - g_ref| / max|g_ref| per tensor (measured on B200, x2 margin).
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,n_out,loss,N", [(256, 198, 4, "xent", None),     # config 3: 256 x 2 s, CE
                                              (64, 298, 64, "ap", 50),        # config 4 wiring: 3 s, AP N=50 D=64
                                              (8, 498, 4, "xent", None)])     # 5 s: T3 = 83 > 56 -> bf16v pooling kernels
def test_training_gradients_bf16_config_sizes(xv, B, T, n_out, loss, N):
    rng = np.random.default_rng(21)
    x = rng.standard_normal((B, T, 40)).astype(np.float32)
    y = rng.integers(0, N or n_out, B)
    params = O.xvector_init(40, n_out, seed=6, bias_scale=0.05)
    m = xv.create((T, 40), n_out, precision="bf16", head="l2_normalize" if loss == "ap" else "log_softmax")
    m.set_weights(params)
    kw = dict(loss="ap", ap_classes=N) if loss == "ap" else {}
    per = m.loss_and_grads(x, y, **kw).cpu().numpy()
    loss_emu, g_emu = _oracle_grads(params, x, y, loss=loss, N=N, emulate_bf16=True)
    assert abs(per.mean() - loss_emu) < 1e-3 * max(1.0, abs(loss_emu))
    _check_grads(m, g_emu, 0.999, GRAD_NW_BOUND)
    if os.environ.get("LBX_TEST_REPORT"):
        grads = m.grads.cpu().numpy()
        for ly in m.layers:
            gw = grads[ly["w_off"]:ly["w_off"] + ly["K"] * ly["ldw"]].reshape(ly["K"], ly["ldw"])[:, :ly["N"]]
            rw = g_emu[ly["name"] + "/kernel"].reshape(ly["K"], ly["N"])
            print("NW", B, T, loss, ly["name"], _nw(gw, rw), _nw(grads[ly["b_off"]:ly["b_off"] + ly["N"]], g_emu[ly["name"] + "/bias"]))
