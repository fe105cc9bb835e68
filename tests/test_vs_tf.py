"""Direct comparison of the oracle with the REAL reference (SURVEY.md §8(c) "Oracle plan").

Runs only where TensorFlow and /root/reference are both present (neither is on the GPU box, and TensorFlow is not
installable in the build container: no wheel, no network), otherwise every test is skipped.  The optional native
dependencies of the reference that are irrelevant to this path (miniaudio, webrtcvad, kaldiio) are stubbed in
sys.modules so that `import lidbox.features.audio` works with TensorFlow alone.

Compared, at the tolerance BASELINE.json:north_star states (1e-4 relative, fp32):
  lidbox/features/audio.py:219-230  spectrograms           lidbox/features/audio.py:247-261  linear_to_mel
  lidbox/features/audio.py:167-174  power_to_db            lidbox/features/audio.py:185-189  ms_to_frames (exact)
  lidbox/models/xvector.py:46-73    create / as_embedding_extractor (weights copied from the Keras model)
  lidbox/losses.py:25-49            SparseAngularProximity
"""
import os
import sys
import types

import numpy as np
import pytest

tf = pytest.importorskip("tensorflow")
REF = os.environ.get("LIDBOX_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REF, "lidbox")):
    pytest.skip("reference checkout not present", allow_module_level=True)

for _name in ("miniaudio", "webrtcvad", "kaldiio"):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import lidbox_oracle as O  # noqa: E402

audio = pytest.importorskip("lidbox.features.audio")


def _signals(B, N, seed=1234):
    rng = np.random.default_rng(seed)
    t = np.arange(N) / 16000.0
    f = rng.uniform(100.0, 4000.0, size=(B, 1))
    return (0.5 * np.sin(2 * np.pi * f * t) + 0.05 * rng.standard_normal((B, N))).astype(np.float32)


def _close(ours, ref, rel=1e-4):
    ref = np.asarray(ref)
    assert ours.shape == ref.shape
    assert np.abs(ours - ref).max() <= rel * max(np.abs(ref).max(), 1e-30)


def test_ms_to_frames_exact():
    for sr in (8000, 16000, 22050, 44100, 48000):
        for ms in range(1, 60):
            assert O.ms_to_frames(sr, ms) == int(audio.ms_to_frames(sr, ms))


@pytest.mark.parametrize("N,fft_length", [(16000, 512), (16399, 512), (8000, 1024), (400, 512)])
def test_spectrograms(N, fft_length):
    x = _signals(3, N)
    _close(O.spectrograms(x, 16000, fft_length=fft_length), audio.spectrograms(x, 16000, fft_length=fft_length).numpy())


def test_linear_to_mel_and_log():
    x = _signals(2, 16000)
    S = audio.spectrograms(x, 16000).numpy()
    M = audio.linear_to_mel(S, 16000).numpy()
    _close(O.linear_to_mel(S, 16000), M)
    from lidbox.features import mel_ops
    W = mel_ops.linear_to_mel_weight_matrix(40, 257, 16000, 0.0, 8000.0).numpy()
    np.testing.assert_allclose(O.linear_to_mel_weight_matrix(40, 257, 16000, 0.0, 8000.0), W, atol=1e-6)
    _close(O.log_eps(M), tf.math.log(M + 1e-6).numpy())


def test_power_to_db():
    S = audio.spectrograms(_signals(2, 8000), 16000).numpy()
    np.testing.assert_allclose(O.power_to_db(S), audio.power_to_db(S).numpy(), atol=2e-3)


def test_xvector_forward_and_embedding():
    xvector = pytest.importorskip("lidbox.models.xvector")
    m = xvector.create((None, 40), 4)
    x = np.random.default_rng(0).standard_normal((3, 98, 40)).astype(np.float32)
    ref = m(x, training=False).numpy()
    params = {}
    for layer in m.layers:
        w = layer.get_weights()
        if len(w) == 2:
            params[layer.name + "/kernel"], params[layer.name + "/bias"] = w
    _close(O.xvector_forward(params, x).astype(np.float32), ref)
    emb = xvector.as_embedding_extractor(m)(x, training=False).numpy()
    _close(O.xvector_forward(params, x, embedding=True).astype(np.float32), emb)


def test_sparse_angular_proximity():
    losses = pytest.importorskip("lidbox.losses")
    rng = np.random.default_rng(1)
    z = rng.standard_normal((6, 16)).astype(np.float32)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    y = rng.integers(0, 10, size=6).astype(np.int32)
    ap = losses.SparseAngularProximity(10, 16, delta_weight=1.5)
    _close(O.ap_loss_per_sample(y, z, 10, 1.5).astype(np.float32), ap.call(y, z).numpy(), rel=1e-5)
    _close(O.ap_theta(z, 10).astype(np.float32), ap.theta(z).numpy(), rel=1e-5)
