"""CPU tests (-m "not gpu"): pin the oracle against everything the reference's own tests hold for this path
 and against the committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import lidbox_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_ms_to_frames_reference_kat():
    # /root/reference/tests/test_features_audio.py:125-129 (exact)
    for sr in range(1000, 60000, 1000):
        for ms in range(1, 5000, 100):
            assert O.ms_to_frames(sr, ms) == (sr // 1000) * ms


def test_ms_to_frames_truncation():
    assert O.ms_to_frames(16000, 25) == 400 and O.ms_to_frames(16000, 10) == 160
    assert O.ms_to_frames(44100, 25) == 1102


def test_fft_frequencies_reference():
    # tests/test_features_audio.py:99-104 compares with librosa.fft_frequencies == np.linspace(0, sr/2, 1+n//2)
    # at 1e-9; the reference returns fp32, so the bound that can hold for non-representable points is fp32 rounding.
    for sr in range(4000, 60000, 4000):
        for n_fft in (2 ** i for i in range(1, 13)):
            a = O.fft_frequencies(sr, n_fft)
            b = np.linspace(0, sr / 2, 1 + n_fft // 2)
            assert a.shape == b.shape
            assert np.abs(a - b).max() <= np.spacing(np.float32(sr / 2))


def test_log10_reference():
    # tests/test_features_audio.py:106-113
    rng = np.random.default_rng(0)
    for rank in range(1, 5):
        x = np.maximum(1e-12, rng.normal(1e6, 1e4, size=rng.integers(1, 10, size=rank)))
        assert np.abs(np.log10(x) - O.log10(x.astype(np.float32))).max() < 1e-6


def test_spectrogram_shapes_reference():
    # tests/test_features_audio.py:131-145: T == N // step - 1 when step = len/2, bins == n_fft//2 + 1
    s = np.random.default_rng(1).standard_normal(48000).astype(np.float32) * 0.1
    for len_ms in range(20, 101, 20):
        for n_fft in (256, 512, 1024, 2048):
            if n_fft < O.ms_to_frames(16000, len_ms):
                continue
            step_ms = len_ms // 2
            P = O.spectrograms(s[None], 16000, frame_length_ms=len_ms, frame_step_ms=step_ms, fft_length=n_fft)[0]
            assert not np.isnan(P).any()
            assert P.shape[0] == s.shape[0] // O.ms_to_frames(16000, step_ms) - 1
            assert P.shape[1] == n_fft // 2 + 1


def test_linear_to_mel_shapes_reference():
    # tests/test_features_audio.py:147-155
    s = np.random.default_rng(2).standard_normal((1, 16000)).astype(np.float32) * 0.1
    P = O.spectrograms(s, 16000)
    for num_mel_bins in range(10, 100, 15):
        M = O.linear_to_mel(P, 16000, num_mel_bins=num_mel_bins)[0]
        assert not np.isnan(M).any()
        assert M.shape == (P.shape[1], num_mel_bins)


def test_power_to_db_reference():
    # tests/test_features_audio.py:115-123
    s = np.random.default_rng(3).standard_normal((1, 16000)).astype(np.float32) * 0.1
    P = O.spectrograms(s, 16000)
    for top_db in range(10, 110, 10):
        db = O.power_to_db(P, top_db=float(top_db))
        assert not np.isnan(db).any()
        assert db.max() <= 0
        assert db.min() >= -top_db - 1e-4


def test_stft_matches_direct_dft():
    # independent check of frame + periodic Hann + end zero-padding against a literal DFT sum
    rng = np.random.default_rng(4)
    x = rng.standard_normal(700)
    L, step, nfft = 400, 160, 512
    S = O.stft(x[None], L, step, nfft, dtype=np.float64)[0]
    assert S.shape == (2, 257)
    n = np.arange(L)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / L)
    for t in range(2):
        fr = x[t * step:t * step + L] * w
        for k in (0, 1, 17, 128, 256):
            ref = (fr * np.exp(-2j * np.pi * k * n / nfft)).sum()
            assert abs(S[t, k] - ref) < 1e-9


def test_mel_matrix_quirks():
    # _linspace divides by num -> FFT bins 241..256 carry no weight, <= 2 non-zeros per bin
    W = O.linear_to_mel_weight_matrix(40, 257, 16000, 0.0, 8000.0)
    assert W.shape == (257, 40) and W.dtype == np.float32
    assert (W[0] == 0).all() and (W[241:] == 0).all()
    assert ((W != 0).sum(axis=1) <= 2).all()
    assert (W != 0).sum() == 464
    assert W.min() >= 0 and W.max() <= 1
    # differs from the true-linspace (TF/HTK) matrix: this oracle is deliberately bug-compatible
    lin = np.linspace(0, 8000, 257)[1:]
    mel = 1127.0 * np.log1p(lin / 700.0)
    e = np.linspace(1127.0 * np.log1p(0 / 700.0), 1127.0 * np.log1p(8000 / 700.0), 42)
    Wt = np.maximum(0, np.minimum((mel[:, None] - e[None, :40]) / (e[None, 1:41] - e[None, :40]),
                                  (e[None, 2:] - mel[:, None]) / (e[None, 2:] - e[None, 1:41])))
    assert np.abs(W[1:] - Wt).max() > 0.5


def test_golden_mel_tables():
    g = np.load(os.path.join(GOLDEN, "mel_tables.npz"))
    for key in g.files:
        _, m, k, sr = key.split("_")
        lo, hi = {"16000": (0.0, 8000.0), "8000": (125.0, 3800.0), "44100": (20.0, 11025.0)}[sr]
        W = O.linear_to_mel_weight_matrix(int(m), int(k), int(sr), lo, hi)
        assert np.array_equal(W, g[key])


def test_golden_wav_fixtures():
    g = np.load(os.path.join(GOLDEN, "wav_fixtures.npz"))
    sig = g["pcm"].astype(np.float32) / np.float32(32768.0)
    lm = O.logmel(sig, 16000, dtype=np.float64)
    assert lm.shape == (5, 48, 40)
    np.testing.assert_allclose(lm, g["logmel"], rtol=1e-5, atol=1e-5)
    lm32 = O.logmel(sig, 16000, dtype=np.float32)
    np.testing.assert_allclose(lm32, g["logmel"], rtol=1e-4, atol=1e-4)


def test_conv_lengths_and_causality():
    rng = np.random.default_rng(5)
    for T in (1, 2, 5, 6, 7, 198):
        x = rng.standard_normal((2, T, 3))
        for k, s in ((5, 1), (3, 2), (3, 3), (1, 1)):
            w = rng.standard_normal((k, 3, 4))
            y = O.conv1d_causal(x, w, np.zeros(4), s, relu=False)
            assert y.shape == (2, -(-T // s), 4)
            # literal definition, SURVEY App. A.8
            t = y.shape[1] - 1
            ref = sum(x[:, t * s + j - (k - 1), :] @ w[j] for j in range(k) if t * s + j - (k - 1) >= 0)
            np.testing.assert_allclose(y[:, t], ref, atol=1e-12)


def test_xvector_edge_shapes_reference():
    # tests/test_models.py:104-107 + testutil.py:29-35: B,T,F >= 1, num_outputs 1..100, no NaN, shape (B, n_out)
    rng = np.random.default_rng(6)
    for (B, T, F, n_out) in ((1, 1, 1, 1), (2, 7, 3, 100), (10, 400, 100, 4), (3, 2, 40, 7)):
        x = rng.uniform(-1e3, 1e3, size=(B, T, F)).astype(np.float32)
        p = O.xvector_init(F, n_out, seed=B)
        y = O.xvector_forward(p, x)
        assert y.shape == (B, n_out) and not np.isnan(y).any()
        np.testing.assert_allclose(np.exp(y.astype(np.float64)).sum(axis=1), 1.0, rtol=1e-4)
    # T == 1 -> std = sqrt(clip) = 1e-5
    x = rng.standard_normal((1, 1, 4)).astype(np.float64)
    _, acts = O.xvector_forward(O.xvector_init(4, 3, dtype=np.float64), x, return_activations=True)
    np.testing.assert_allclose(acts["stats_pooling"][0, 1500:], 1e-5)


def test_xvector_numpy_vs_torch_twin_and_golden():
    import torch
    g = np.load(os.path.join(GOLDEN, "xvector_small.npz"))
    params = O.xvector_init(24, 5, seed=11, bias_scale=0.05)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    logp = O.xvector_forward(p64, g["x"].astype(np.float64))
    np.testing.assert_allclose(logp, g["logp"], atol=1e-10)
    np.testing.assert_allclose(O.xvector_forward(p64, g["x"].astype(np.float64), embedding=True), g["emb"], atol=1e-10)
    tp = {k: torch.tensor(v, dtype=torch.float64) for k, v in p64.items()}
    lp_t = O.torch_xvector_forward(tp, torch.tensor(g["x"], dtype=torch.float64)).numpy()
    np.testing.assert_allclose(lp_t, logp, atol=1e-10)
    assert abs(O.sparse_xent_on_logprobs(g["y"], logp) - float(g["loss"])) < 1e-12


def test_param_count_reference():
    # 4 508 124 + 513 * num_outputs parameters for F = 40
    for n_out in (4, 50):
        n = sum(int(np.prod(s)) for s in O.xvector_param_shapes(40, n_out).values())
        assert n == 4508124 + 513 * n_out


def test_ap_loss_golden_and_properties():
    g = np.load(os.path.join(GOLDEN, "ap_loss.npz"))
    N, w = int(g["N"]), float(g["delta_weight"])
    per = O.ap_loss_per_sample(g["y"], g["z"], N, w)
    np.testing.assert_allclose(per, g["per_sample"], atol=1e-12)
    assert abs(per.mean() - float(g["loss"])) < 1e-12
    # losses.py:48-49 with one-hot c_T: theta == acos of the first N coordinates
    c_T = np.eye(N, g["z"].shape[1]).T
    np.testing.assert_allclose(O.ap_theta(g["z"], N), np.arccos(g["z"] @ c_T), atol=1e-12)
    # gradient is zero for coordinates >= N
    assert np.abs(g["grad"][:, N:]).max() == 0
    # a perfectly aligned vector has the smallest loss for its class
    z = np.eye(12)[:3] * 0.999
    assert O.ap_loss_per_sample([0, 1, 2], z, N).max() < O.ap_loss_per_sample([1, 2, 0], z, N).min()


def test_torch_logmel_baseline_matches_oracle():
    import torch
    s = np.random.default_rng(8).standard_normal((2, 16000)).astype(np.float32) * 0.1
    a = O.torch_logmel(torch.from_numpy(s)).numpy()
    b = O.logmel(s, 16000, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-4)


def test_feature_scaling_reference():
    # /root/reference/tests/test_features.py:14-26 (fewer repetitions)
    rng = np.random.default_rng(10)
    for rank in range(1, 5):
        for _ in range(20):
            delta = rng.uniform(1, 1e3)
            lo = rng.uniform(-delta, delta)
            hi = lo + rng.uniform(0, delta / 2)
            x = rng.normal(0, delta ** 2, size=rng.integers(2, 20, size=rank))
            for axis in [None] + list(range(rank)):
                y = O.feature_scaling(x, lo, hi, axis=axis)
                assert not np.isnan(y).any() and y.shape == x.shape
                assert np.abs(y.min(axis=axis) - lo).max() < 1e-9
                assert np.abs(y.max(axis=axis) - hi).max() < 1e-9


def test_cmvn_reference():
    # tests/test_features.py:28-43
    rng = np.random.default_rng(11)
    for delta_magnitude in range(2, 7):
        for _ in range(20):
            delta = rng.uniform(1, 10 ** delta_magnitude)
            x = rng.uniform(-delta, delta, size=rng.integers(1, 20, size=3))
            for axis in range(3):
                y_m = O.cmn(x, axis=axis)
                assert not np.isnan(y_m).any() and y_m.shape == x.shape
                assert np.abs(y_m.mean(axis=axis)).max() < 1
                y_mv = O.cmvn(x, axis=axis)
                assert not np.isnan(y_mv).any() and y_mv.shape == x.shape
                assert np.abs(y_mv.mean(axis=axis)).max() < 0.1
                assert y_mv.var(axis=axis).max() < 10


def test_window_normalization_reference():
    # tests/test_features.py:45-58 + the reference's own NumPy variant (features/__init__.py:90-110) for odd windows,
    # whose clipped (un-padded) windows coincide with the reflected ones in the interior of the signal
    rng = np.random.default_rng(12)
    for _ in range(20):
        x = rng.uniform(-100, 100, size=rng.integers(1, 20, size=3))
        for window_len in [-1] + list(range(2, x.shape[0] + 1)):
            for nv in (True, False):
                y = O.window_normalization(x, axis=1, window_len=window_len, normalize_variance=nv)
                assert not np.isnan(y).any() and y.shape == x.shape
    x = rng.standard_normal((2, 50, 3))
    w = 7
    y = O.window_normalization(x, window_len=w, normalize_variance=True)
    for t in range(w // 2, 50 - w // 2):
        win = x[:, t - w // 2:t - w // 2 + w]
        np.testing.assert_allclose(y[:, t], (x[:, t] - win.mean(axis=1)) / win.std(axis=1), atol=1e-12)


def test_run_length_encoding_reference_kat():
    # /root/reference/tests/test_features_audio.py:166-169 (exact)
    pos, length = O.run_length_encoding(np.array([1, 1, 1, 2, 2, 2, 3, 4, 5, 6, 6, 7]))
    assert (pos == np.array([0, 3, 6, 7, 8, 9, 11])).all()
    assert (length == np.array([3, 3, 1, 1, 1, 2, 1])).all()


def test_root_mean_square_reference():
    # tests/test_features_audio.py:157-163
    rng = np.random.default_rng(13)
    for _ in range(20):
        x = rng.normal(0, 5, size=rng.integers(1, 10, size=2))
        assert np.abs(np.sqrt(np.mean(np.square(np.abs(x)), axis=-1)) - O.root_mean_square(x, axis=-1)).max() < 1e-5


def test_vad_reference_properties():
    # tests/test_features_audio.py:175-191 on the reference's WAV fixtures (first 0.5 s, carried by the golden file)
    g = np.load(os.path.join(GOLDEN, "wav_fixtures.npz"))
    for pcm in g["pcm"]:
        s = pcm.astype(np.float32) / np.float32(32768.0)
        assert O.framewise_rms_energy_vad_decisions(s, 16000, 25).all()
        assert O.remove_silence(s, 16000).shape == s.shape
    z = np.zeros(3 * 16000, np.float32)
    assert not O.framewise_rms_energy_vad_decisions(z, 16000, 25).any()
    assert O.remove_silence(z, 16000).size == 0
    # too-short non-speech runs are flipped back to speech
    m = np.array([1, 0, 0, 1, 0, 0, 0, 0, 1, 0], bool)
    assert (O.invert_too_short_consecutive_false(m, 3) == np.array([1, 1, 1, 1, 0, 0, 0, 0, 1, 1], bool)).all()
    assert (O.invert_too_short_consecutive_false(m, 0) == m).all()


def test_mfcc_is_orthonormal_dct():
    # tf.signal.mfccs_from_log_mel_spectrograms == scipy's orthonormal DCT-II except for the k = 0 scale (sqrt(2))
    import scipy.fft
    x = np.random.default_rng(14).standard_normal((3, 7, 40))
    m = O.mfccs_from_log_mel_spectrograms(x)
    ref = scipy.fft.dct(x, type=2, norm="ortho", axis=-1)
    np.testing.assert_allclose(m[..., 1:], ref[..., 1:], atol=1e-12)
    np.testing.assert_allclose(m[..., 0], ref[..., 0] * np.sqrt(2.0), atol=1e-12)


def test_xvector_extended_reference_shapes():
    import torch
    # /root/reference/lidbox/models/xvector_extended.py:25-40 + tests/test_models.py:104-107 (shape / no-NaN checks)
    ext = dict(frame_layers=O.XVECTOR_EXTENDED_FRAME_LAYERS, output_name="output")
    shapes = O.xvector_param_shapes(40, 7, **ext)
    assert shapes["frame7/kernel"] == (3, 512, 512) and shapes["frame10/kernel"] == (1, 512, 1500)
    assert shapes["segment1/kernel"] == (3000, 512) and shapes["output/kernel"] == (512, 7)
    n = sum(int(np.prod(s)) for s in shapes.values())
    conv = 5 * 40 * 512 + 512 * 512 + 3 * 512 * 512 + 512 * 512 + 3 * 512 * 512 + 512 * 512 + 3 * 512 * 512 \
        + 512 * 512 + 512 * 512 + 512 * 1500 + 9 * 512 + 1500
    assert n == conv + 3000 * 512 + 512 + 512 * 512 + 512 + 512 * 7 + 7
    rng = np.random.default_rng(0)
    for (B, T, F, n_out) in ((1, 1, 1, 1), (3, 50, 24, 5), (2, 400, 100, 100)):
        params = O.xvector_init(F, n_out, seed=1, bias_scale=0.05, **ext)
        x = rng.standard_normal((B, T, F))
        p64 = {k: v.astype(np.float64) for k, v in params.items()}
        lp, acts = O.xvector_forward(p64, x, return_activations=True, **ext)
        assert lp.shape == (B, n_out) and not np.isnan(lp).any()
        np.testing.assert_allclose(np.exp(lp).sum(axis=1), 1.0, rtol=1e-9)
        assert acts["frame10"].shape == (B, -(-(-(-(-(-T // 2)) // 3)) // 4), 1500)     # ceil(ceil(ceil(T/2)/3)/4)
        tp = {k: torch.tensor(v) for k, v in p64.items()}
        lp_t = O.torch_xvector_forward(tp, torch.tensor(x), **ext).numpy()
        np.testing.assert_allclose(lp_t, lp, rtol=1e-9, atol=1e-9)
        raw = O.xvector_forward(p64, x, output_activation=None, **ext)
        np.testing.assert_allclose(raw - np.log(np.exp(raw).sum(axis=1, keepdims=True)), lp, rtol=1e-9, atol=1e-9)


def test_activation_row_geometry_for_any_layer_list():
    # host logic of models/xvector.py: consecutive buffers share one row geometry, also when kernel_size < strides
    from lidbox_b200.models.xvector import _Geometry, frame_layer
    from lidbox_b200.models.xvector_extended import frame_layers
    lists = [frame_layers(), [frame_layer(8, 5, 1), frame_layer(8, 3, 2), frame_layer(8, 3, 3), frame_layer(8, 1, 1)],
             [frame_layer(8, 3, 4)], [frame_layer(8, 2, 5), frame_layer(8, 7, 2)],
             # dilated TDNN (extension): causal padding dilation * (k - 1)
             [frame_layer(8, 5, 1), frame_layer(8, 3, 1, dilation_rate=2), frame_layer(8, 3, 1, dilation_rate=3),
              frame_layer(8, 1, 1)],
             [frame_layer(8, 3, 2), frame_layer(8, 5, 1, dilation_rate=4)]]
    for frames in lists:
        for T in (1, 2, 5, 23, 24, 25, 198, 400, 499):
            geo = _Geometry(T, frames)
            for L, f in enumerate(frames):
                assert geo.T[L + 1] == -(-geo.T[L] // f.strides)
                pad = (f.kernel_size - 1) * f.dilation_rate
                assert geo.pad[L] == pad and geo.Tpad[L] >= geo.T[L] + pad   # room for the causal left padding
                assert geo.Tpad[L] % f.strides == 0 and geo.R[L] >= geo.T[L + 1]
                if L + 1 < len(frames):
                    assert geo.R[L] == geo.Tpad[L + 1]


# --------------------------------------------------------------------------- chunking / merging / C_avg
_CAVG_TRUE = np.array([[1, 0, 0], [0, 1, 0], [0, 1, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 0, 1]],
                      np.float32)                               # the self-test inputs of lidbox/metrics.py:128-150
_CAVG_PROB = np.array([[.1, .2, .9], [.9, .2, .0], [.1, .9, .0], [.2, .8, .5], [.6, .3, .1], [.1, .0, .7],
                       [.1, .0, .7], [.9, .1, .0]], np.float32)


def test_cavg_reference_selftest_and_definition():
    with np.errstate(divide="ignore"):
        pred = np.log(_CAVG_PROB)
    thresholds = np.log(np.array([0.05, 0.4, 0.6, 0.95], np.float32))
    cavg = O.AverageDetectionCost(3, thresholds)
    assert cavg.result() == 0.0
    cavg.update_state(_CAVG_TRUE, pred)
    labels = _CAVG_TRUE.argmax(axis=1)
    per_t = cavg.result_per_threshold()
    want = [O.cavg_by_definition(labels, pred, t, 3) for t in thresholds]
    np.testing.assert_allclose(per_t, want, rtol=1e-6)
    assert cavg.result() == per_t.min() and 0.0 < cavg.result() < 0.5
    # metrics.py:113-119 (_assert_P_fa): the l == m pairs never receive a count
    idx = np.arange(3)
    assert not cavg.fp_pairs[idx, idx].any() and not cavg.tn_pairs[idx, idx].any()
    # counters are trial counts: every trial lands in exactly one of tp/fn per threshold
    np.testing.assert_array_equal((cavg.tp + cavg.fn).sum(axis=0), len(labels))
    np.testing.assert_array_equal((cavg.fp_pairs + cavg.tn_pairs).sum(axis=(0, 1)), len(labels) * 2)
    # metrics.py:160-161: result is 0 after reset_states
    cavg.reset_states()
    assert cavg.result() == 0.0
    # sparse variant == dense variant; two half batches == one batch
    a, b = O.SparseAverageDetectionCost(3, thresholds), O.AverageDetectionCost(3, thresholds)
    a.update_state(labels[:5], pred[:5])
    a.update_state(labels[5:], pred[5:])
    b.update_state(_CAVG_TRUE, pred)
    assert a.result() == b.result()
    # a perfect system has zero cost at a separating threshold, an inverted one the maximum P_tar*C_miss + (1-P_tar)*C_fa
    perfect = np.where(_CAVG_TRUE > 0, 0.0, -10.0)
    c = O.AverageDetectionCost(3, [-5.0])
    c.update_state(_CAVG_TRUE, perfect)
    assert c.result() == 0.0
    c = O.AverageDetectionCost(3, [-5.0])
    c.update_state(_CAVG_TRUE, -10.0 - perfect)
    assert c.result() == 1.0


def test_create_signal_chunks_reference_semantics():
    rng = np.random.default_rng(0)
    for (n, sr, length_ms, step_ms, pad_ms) in ((16000, 16000, 500, 250, 0), (16001, 16000, 500, 250, 0),
                                                (20000, 16000, 1000, 1000, 0), (20000, 16000, 1000, 1000, 800),
                                                (20000, 16000, 1000, 1000, 700), (100, 16000, 500, 250, 0),
                                                (100, 16000, 500, 250, 500), (7999, 8000, 2000, 500, 10),
                                                (48000, 44100, 30, 10, 5)):
        sig = rng.standard_normal(n).astype(np.float32)
        chunks = O.create_signal_chunks(sig, sr, length_ms, step_ms, pad_ms)
        L = int(np.int32(np.float32(sr) * np.float32(1e-3 * length_ms)))
        step = int(np.int32(np.float32(sr) * np.float32(1e-3 * step_ms)))
        pad = int(np.int32(np.float32(sr) * np.float32(1e-3 * pad_ms)))
        full = max(0, 1 + (n - L) // step)
        assert chunks.shape[1] == L and chunks.shape[0] in (full, full + 1)
        for c in range(full):
            np.testing.assert_array_equal(chunks[c], sig[c * step:c * step + L])
        if chunks.shape[0] == full + 1:               # padded tail chunk: the rest of the signal, then zeros
            rest = n - full * step
            assert rest < L <= rest + pad
            np.testing.assert_array_equal(chunks[full, :rest], sig[full * step:])
            assert not chunks[full, rest:].any()
    assert O.create_signal_chunks(np.zeros(20000, np.float32), 16000, 1000, 1000, 800).shape == (2, 16000)
    assert O.create_signal_chunks(np.zeros(20000, np.float32), 16000, 1000, 1000, 700).shape == (1, 16000)
    assert O.create_signal_chunks(np.zeros(100, np.float32), 16000, 500, 250, 0).shape == (0, 8000)


def test_merge_chunk_predictions_reference():
    ids = ["utt-b-000001", "utt-a-000002", "utt-b-000002", "utt-a-000001", "solo-000001"]
    pred = np.arange(15, dtype=np.float32).reshape(5, 3)
    parents, merged = O.merge_chunk_predictions(ids, pred)
    assert parents == ["solo", "utt-a", "utt-b"]
    np.testing.assert_allclose(merged, [pred[4], (pred[1] + pred[3]) / 2, (pred[0] + pred[2]) / 2])


def test_chunk_host_logic_matches_oracle():
    # host side of lidbox_b200/data/steps.py (no GPU): geometry, chunk counts through the C-ABI, chunk ids
    from lidbox_b200.data import steps
    for (n, sr, length_ms, step_ms, pad_ms) in ((16000, 16000, 500, 250, 0), (20000, 16000, 1000, 1000, 800),
                                                (20000, 16000, 1000, 1000, 700), (100, 16000, 500, 250, 0),
                                                (100, 16000, 500, 250, 500), (7999, 8000, 2000, 500, 10),
                                                (48000, 44100, 30, 10, 5), (0, 16000, 500, 250, 500)):
        L, step, pad = steps.chunk_geometry(sr, length_ms, step_ms, pad_ms)
        want = O.create_signal_chunks(np.zeros(n, np.float32), sr, length_ms, step_ms, pad_ms)
        assert (steps.num_signal_chunks(n, L, step, pad), L) == want.shape
    assert steps.chunk_ids("utt", 3) == ["utt-000001", "utt-000002", "utt-000003"]     # steps.py:589-591: width 6
    assert steps.chunk_ids("utt", 1, max_num_chunks_per_signal=100) == ["utt-01"]
    with pytest.raises(ValueError):
        steps.chunk_geometry(16000, 0, 10)


def test_new_golden_fixtures_match_oracle():
    # tests/golden/xvector_extended_small.npz and cavg.npz (make_golden.py sections 5-6) pin the restatement
    g = np.load(os.path.join(GOLDEN, "xvector_extended_small.npz"))
    ext = dict(frame_layers=O.XVECTOR_EXTENDED_FRAME_LAYERS, output_name="output")
    pe = {k: v.astype(np.float64) for k, v in O.xvector_init(24, 5, seed=12, bias_scale=0.05, **ext).items()}
    np.testing.assert_allclose(O.xvector_forward(pe, g["x"].astype(np.float64), **ext), g["logp"], rtol=1e-10, atol=1e-12)
    c = np.load(os.path.join(GOLDEN, "cavg.npz"))
    m = O.AverageDetectionCost(3, c["thr"])
    m.update_state(c["onehot"], c["pred"])
    np.testing.assert_array_equal(m.fp_pairs, c["fp_pairs"])
    np.testing.assert_allclose(m.result_per_threshold(), c["cavg"], rtol=1e-7)
    m6 = O.SparseAverageDetectionCost(6, c["thr6"], C_miss=1.0, C_fa=2.0, P_tar=0.3)
    m6.update_state(c["y6"], c["s6"])
    np.testing.assert_allclose(m6.result_per_threshold(), c["cavg6"], rtol=1e-7)
    want = [O.cavg_by_definition(c["y6"], c["s6"], t, 6, 1.0, 2.0, 0.3) for t in c["thr6"]]
    np.testing.assert_allclose(c["cavg6"], want, rtol=1e-5)
