"""
CPU ORACLE — test infrastructure only, never part of the product path.

A bug-compatible CPU restatement (NumPy, plus torch-CPU autograd for gradients)
of the lidbox hot path named by BASELINE.json:north_star.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; ``lidbox_b200`` never does.

PARITY PINNING STATUS
---------------------
The arithmetic of this path lives in a third-party dependency that is absent from
/root/reference: **TensorFlow ~= 2.3.0** (requirements-test.txt:4).  TensorFlow is
not installable in this image (no wheel, no network), so the reference cannot be
run here.  What IS pinned by the reference's own tests and is checked in
tests/test_oracle.py:
  * ms_to_frames         exact KAT (tests/test_features_audio.py:125-129)
  * frame count / bins   T == N // step - 1 for step = len/2 (tests/test_features_audio.py:131-145)
  * fft_frequencies      vs linspace 1e-9   (tests/test_features_audio.py:99-104)
  * log10                vs numpy 1e-6      (tests/test_features_audio.py:106-113)
  * power_to_db          max <= 0, no NaN   (tests/test_features_audio.py:115-123)
  * x-vector output      shape / no NaN for B,T,F >= 1 (tests/test_models.py:104-107)
For spectrogram / mel / log-mel VALUES, x-vector outputs, gradients and the AP
loss the reference holds no golden vectors, and no number produced by TensorFlow
exists in this repository: against the reference itself those values stay
**parity unpinned**.  What closes most of the gap (round 2):
  * tests/test_third_party.py checks every restated TF semantic against
    independent third-party code in the image — scipy.signal.stft / ShortTimeFFT and
    transformers.audio_utils (framing, Hann, END zero-padding, rFFT), transformers'
    HTK mel filter bank (formula family; the reference's two off-by-one divisors are
    literal in mel_ops.py:16,40-41,49-55), scipy.fft.dct (MFCC), torch conv1d /
    linear / log_softmax / var (TDNN), torch autograd + finite differences (AP loss),
    torch.optim.Adam (optimizer algebra).  TF's odd-length "periodic" Hann
    (n = L + periodic*even - 1 -> symmetric window) follows the in-tree copy of the
    TF formula in blackman_window (lidbox/features/audio.py:206-211).
  * tests/test_vs_tf.py compares directly with the real reference and runs wherever
    TensorFlow + /root/reference exist (skipped here).
Still restated from memory of TF/Keras 2.3 with no executable confirmation: Keras
Conv1D padding="causal" = left pad of k-1 and stride phase t*s + j; Keras Adam's
epsilon placement (outside the bias-corrected sqrt); glorot-uniform limits.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
import numpy as np

# --------------------------------------------------------------------------- #
# lidbox/features/audio.py
# --------------------------------------------------------------------------- #


def ms_to_frames(sample_rate, ms):
    """lidbox/features/audio.py:185-189 — int32(float32(sr) * 1e-3f * float32(ms)), left to right in fp32."""
    v = np.float32(sample_rate) * np.float32(1e-3)
    v = np.float32(v) * np.float32(ms)
    return int(np.int32(v))


def fft_frequencies(sample_rate, n_fft):
    """lidbox/features/audio.py:150-159 — tf.linspace(0, sr//2, 1 + n_fft//2) in fp32."""
    return np.linspace(0.0, float(sample_rate // 2), 1 + n_fft // 2).astype(np.float32)


def log10(x):
    """lidbox/features/audio.py:162-164 — ln(x) / ln(10)."""
    x = np.asarray(x)
    return np.log(x) / np.log(np.asarray(10.0, x.dtype))


def hann_window(length, dtype=np.float32):
    """tf.signal.hann_window(L, periodic=True) (call site audio.py:229 through tf.signal.stft).
    even = 1 - L % 2 ; n = L + even - 1 ; w[i] = 0.5 - 0.5 cos(2 pi i / n)  (SURVEY App. A.2;
    same formula family as the in-tree blackman_window, audio.py:192-216)."""
    if length == 1:
        return np.ones(1, dtype)
    even = 1 - length % 2
    n = dtype(length + even - 1)
    i = np.arange(length, dtype=dtype)
    arg = dtype(2.0 * np.pi) * i / n
    return (dtype(0.5) - dtype(0.5) * np.cos(arg)).astype(dtype)


def num_frames(n_samples, frame_length, frame_step):
    """tf.signal.frame(pad_end=False): T = max(0, 1 + (N - L) // step)."""
    if n_samples < frame_length:
        return 0
    return 1 + (n_samples - frame_length) // frame_step


def stft(signals, frame_length, frame_step, fft_length, dtype=np.float32):
    """tf.signal.stft(signals, L, step, fft_length) (call site audio.py:229):
    frame (no centre/end padding) -> periodic Hann -> zero-pad at the END to fft_length -> rFFT."""
    signals = np.asarray(signals, dtype)
    assert signals.ndim == 2
    B, N = signals.shape
    T = num_frames(N, frame_length, frame_step)
    K = fft_length // 2 + 1
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    if T == 0:
        return np.zeros((B, 0, K), cdtype)
    idx = np.arange(T)[:, None] * frame_step + np.arange(frame_length)[None, :]
    frames = signals[:, idx] * hann_window(frame_length, dtype)[None, None, :]
    # numpy's pocketfft computes in the input precision for float32 input via scipy only;
    # np.fft upcasts to double, which is the more accurate "truth" – we round at the end.
    S = np.fft.rfft(frames.astype(np.float64), n=fft_length, axis=-1)
    return S.astype(cdtype)


def spectrograms(signals, sample_rate, frame_length_ms=25, frame_step_ms=10, power=2.0, fft_length=512,
                 dtype=np.float32):
    """lidbox/features/audio.py:219-230 — pow(abs(stft), power)."""
    frame_length = ms_to_frames(sample_rate, frame_length_ms)
    frame_step = ms_to_frames(sample_rate, frame_step_ms)
    S = stft(signals, frame_length, frame_step, fft_length, dtype)
    return np.power(np.abs(S).astype(dtype), dtype(power)).astype(dtype)


# --------------------------------------------------------------------------- #
# lidbox/features/mel_ops.py
# --------------------------------------------------------------------------- #

_MEL_BREAK_FREQUENCY_HERTZ = 700.0
_MEL_HIGH_FREQUENCY_Q = 1127.0


def _linspace(start, stop, num, dtype=np.float32):
    """lidbox/features/mel_ops.py:11-16 — start + (stop - start) * range / num   (divides by num, NOT num-1)."""
    rng = np.arange(0, num, dtype=dtype)
    start = dtype(start)
    stop = dtype(stop)
    return (start + (stop - start) * rng / dtype(num)).astype(dtype)


def _hertz_to_mel(f, dtype=np.float32):
    """lidbox/features/mel_ops.py:23-25."""
    f = np.asarray(f, dtype)
    return (dtype(_MEL_HIGH_FREQUENCY_Q) * np.log(dtype(1.0) + f / dtype(_MEL_BREAK_FREQUENCY_HERTZ))).astype(dtype)


def linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000,
                                lower_edge_hertz=125.0, upper_edge_hertz=3800.0, dtype=np.float32):
    """lidbox/features/mel_ops.py:28-75, same evaluation order, all in `dtype` (reference: fp32)."""
    bands_to_zero = 1
    nyquist = dtype(sample_rate) / dtype(2.0)
    linear_frequencies = _linspace(0.0, nyquist, num_spectrogram_bins, dtype)[bands_to_zero:]
    spectrogram_bins_mel = _hertz_to_mel(linear_frequencies, dtype)[:, None]
    edges = _linspace(_hertz_to_mel(lower_edge_hertz, dtype), _hertz_to_mel(upper_edge_hertz, dtype),
                      num_mel_bins + 2, dtype)
    # tf.signal.frame(edges, 3, 1): num_mel_bins triples (lower, center, upper)
    lower = edges[0:num_mel_bins][None, :]
    center = edges[1:num_mel_bins + 1][None, :]
    upper = edges[2:num_mel_bins + 2][None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        lower_slopes = (spectrogram_bins_mel - lower) / (center - lower)
        upper_slopes = (upper - spectrogram_bins_mel) / (upper - center)
        W = np.maximum(dtype(0.0), np.minimum(lower_slopes, upper_slopes))
    return np.pad(W, [[bands_to_zero, 0], [0, 0]]).astype(dtype)


def linear_to_mel(S, sample_rate, num_mel_bins=40, fmin=0.0, fmax=8000.0, dtype=np.float32):
    """lidbox/features/audio.py:247-261 — tensordot(S, W, 1)."""
    S = np.asarray(S, dtype)
    W = linear_to_mel_weight_matrix(num_mel_bins, S.shape[2], sample_rate, fmin, fmax, np.float32).astype(dtype)
    return np.tensordot(S, W, 1).astype(dtype)


def log_eps(X, eps=1e-6):
    """lidbox/data/tf_utils.py:178 — ln(X + 1e-6)."""
    X = np.asarray(X)
    return np.log(X + X.dtype.type(eps))


def power_to_db(S, amin=1e-10, top_db=80.0):
    """lidbox/features/audio.py:167-174 — 20*(log10(max(amin,S)) - log10(max(amin,max_all(S)))), floored at max-top_db."""
    S = np.asarray(S, np.float32)
    amin = np.float32(amin)
    db = np.float32(20.0) * (log10(np.maximum(amin, S)) - log10(np.maximum(amin, S.max())))
    return np.maximum(db, db.max() - np.float32(top_db)).astype(np.float32)


def db_to_power(S):
    """lidbox/features/audio.py:177-181 — pow(10, S / 20)."""
    S = np.asarray(S)
    return np.power(S.dtype.type(10.0), S / S.dtype.type(20.0))


def logmel(signals, sample_rate, frame_length_ms=25, frame_step_ms=10, power=2.0, fft_length=512,
           num_mel_bins=40, fmin=0.0, fmax=8000.0, dtype=np.float32):
    """The intended map-stage chain (tf_utils.py:172-178): spectrograms -> linear_to_mel -> ln(x+1e-6)."""
    S = spectrograms(signals, sample_rate, frame_length_ms, frame_step_ms, power, fft_length, dtype)
    M = linear_to_mel(S, sample_rate, num_mel_bins, fmin, fmax, dtype)
    return log_eps(M)


def mfccs_from_log_mel_spectrograms(log_mel):
    """tf.signal.mfccs_from_log_mel_spectrograms (call site tf_utils.py:183): DCT-II (tf.signal.dct type 2, no norm:
    X_k = 2 sum_n x_n cos(pi k (2n+1) / (2N))) scaled by rsqrt(2 N)."""
    x = np.asarray(log_mel, np.float64)
    N = x.shape[-1]
    n = np.arange(N)
    C = 2.0 * np.cos(np.pi * np.arange(N)[:, None] * (2 * n[None, :] + 1) / (2.0 * N))
    return (x @ C.T) / np.sqrt(2.0 * N)


def extract_features(signals, sample_rates, feattype, spec_kwargs=None, melspec_kwargs=None, mfcc_kwargs=None,
                     db_spec_kwargs=None, feat_scale_kwargs=None, window_norm_kwargs=None):
    """lidbox/data/tf_utils.py:166-195 (with the broken `melspectrograms` name read as `linear_to_mel`).
    """
    signals = np.asarray(signals, np.float32)
    if signals.ndim != 2:
        raise ValueError("signals must be [B, N]")
    sample_rates = np.asarray(sample_rates)
    if not (sample_rates == sample_rates[0]).all():
        raise ValueError("different sample rates in a batch")
    sr = int(sample_rates[0])
    X = spectrograms(signals, sr, **(spec_kwargs or {}))
    if feattype in ("melspectrogram", "logmelspectrogram", "mfcc"):
        X = linear_to_mel(X, sr, **(melspec_kwargs or {}))
        if feattype in ("logmelspectrogram", "mfcc"):
            X = log_eps(X)
        if feattype == "mfcc":
            mk = mfcc_kwargs or {}
            X = mfccs_from_log_mel_spectrograms(X)[..., mk.get("coef_begin", 1):mk.get("coef_end", 13)]
    elif feattype == "db_spectrogram":
        X = power_to_db(X, **(db_spec_kwargs or {}))
    # any other feattype string (incl. "spectrogram") falls through the reference's if/elif chain: X stays the
    # power spectrogram (tf_utils.py:172-188)
    if not np.isfinite(X).all():
        raise FloatingPointError(feattype + " failed")
    if feat_scale_kwargs:
        X = feature_scaling(X, **feat_scale_kwargs)           # tf_utils.py:189-191
    if window_norm_kwargs:
        X = window_normalization(X, **window_norm_kwargs)     # tf_utils.py:192-194
    return X


# --------------------------------------------------------------------------- #
# energy VAD, lidbox/features/audio.py:262-353  
# --------------------------------------------------------------------------- #


def root_mean_square(x, axis=-1):
    """lidbox/features/audio.py:265-269."""
    x = np.asarray(x, np.float64)
    return np.sqrt(np.mean(np.square(np.abs(x)), axis=axis))


def run_length_encoding(v):
    """lidbox/features/audio.py:276-283."""
    v = np.asarray(v).reshape(-1)
    i = np.concatenate(([-1], np.flatnonzero(v[1:] != v[:-1]), [v.size - 1]))
    pos = np.concatenate(([0], np.cumsum(i[1:] - i[:-1])))
    return pos[:-1], pos[1:] - pos[:-1]


def invert_too_short_consecutive_false(mask, min_length):
    """lidbox/features/audio.py:289-296."""
    mask = np.asarray(mask, bool)
    if min_length == 0:
        return mask
    pos, lengths = run_length_encoding(mask.astype(np.int32))
    return np.repeat(np.logical_or(mask[pos], lengths < min_length), lengths)


def framewise_rms_energy_vad_decisions(signal, sample_rate, frame_step_ms, min_non_speech_ms=0, strength=0.05,
                                       min_rms_threshold=1e-3, return_margin=False):
    """lidbox/features/audio.py:307-329 (non-overlapping frames of ms_to_frames(sr, frame_step_ms) samples)."""
    signal = np.asarray(signal, np.float64)
    step = ms_to_frames(sample_rate, frame_step_ms)
    F = signal.shape[0] // step
    frames = signal[:F * step].reshape(F, step)
    rms = root_mean_square(frames, axis=1)
    threshold = strength * max(min_rms_threshold, rms.mean()) if F else 0.0
    decisions = rms > threshold
    min_frames = int(ms_to_frames(sample_rate, min_non_speech_ms) / step)
    out = invert_too_short_consecutive_false(decisions, min_frames)
    if return_margin:       # relative distance of every frame's RMS from the threshold (ties are not comparable in fp32)
        return out, np.abs(rms - threshold) / max(threshold, 1e-30)
    return out


def remove_silence(signal, rate, window_ms=10, min_non_speech_ms=300):
    """lidbox/features/audio.py:337-353."""
    signal = np.asarray(signal)
    window = (window_ms * rate) // 1000
    vad = framewise_rms_energy_vad_decisions(signal, rate, window_ms, min_non_speech_ms=min_non_speech_ms, strength=0.1)
    F = signal.shape[0] // window
    return signal[:F * window].reshape(F, window)[vad[:F]].reshape(-1)


# --------------------------------------------------------------------------- #
# lidbox/features/__init__.py  (feature normalisation)
# --------------------------------------------------------------------------- #


def _divide_no_nan(a, b):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(b == 0, 0.0, a / np.where(b == 0, 1.0, b))


def feature_scaling(X, min, max, axis=None):
    """lidbox/features/__init__.py:5-9."""
    X = np.asarray(X, np.float64)
    x_min = X.min(axis=axis, keepdims=True)
    x_max = X.max(axis=axis, keepdims=True)
    return min + (max - min) * _divide_no_nan(X - x_min, x_max - x_min)


def cmn(X, axis=1):
    """lidbox/features/__init__.py:16-20."""
    X = np.asarray(X, np.float64)
    return X - X.mean(axis=axis, keepdims=True)


def cmvn(X, axis=1):
    """lidbox/features/__init__.py:26-32 (tf.math.reduce_std = population standard deviation)."""
    X = np.asarray(X, np.float64)
    return _divide_no_nan(cmn(X, axis), X.std(axis=axis, keepdims=True))


def window_normalization(X, axis=1, window_len=-1, normalize_variance=True):
    """lidbox/features/__init__.py:40-67: REFLECT padding (w//2 left, w//2 - 1 + (w & 1) right), tf.signal.frame(step 1)."""
    X = np.asarray(X, np.float64)
    if window_len == -1 or X.shape[1] <= window_len:
        return cmvn(X, axis) if normalize_variance else cmn(X, axis)
    assert axis == 1
    w = window_len
    Xp = np.pad(X, [(0, 0), (w // 2, w // 2 - 1 + (w & 1)), (0, 0)], mode="reflect")
    idx = np.arange(X.shape[1])[:, None] + np.arange(w)[None, :]
    windows = Xp[:, idx, :]                                   # [B, T, w, F]
    out = X - windows.mean(axis=2)
    if normalize_variance:
        out = _divide_no_nan(out, windows.std(axis=2))
    return out


# --------------------------------------------------------------------------- #
# lidbox/models/xvector.py  (NumPy forward; torch-CPU twin below for gradients)
# --------------------------------------------------------------------------- #

STDDEV_SQRT_MIN_CLIP = 1e-10     # lidbox/models/xvector.py:22
FRAME_LAYERS = (                  # lidbox/models/xvector.py:53-57  (filters, kernel_size, strides)
    ("frame1", 512, 5, 1),
    ("frame2", 512, 3, 2),
    ("frame3", 512, 3, 3),
    ("frame4", 512, 1, 1),
    ("frame5", 1500, 1, 1),
)
XVECTOR_EXTENDED_FRAME_LAYERS = (  # lidbox/models/xvector_extended.py:25-34
    ("frame1", 512, 5, 1),
    ("frame2", 512, 1, 1),
    ("frame3", 512, 3, 2),
    ("frame4", 512, 1, 1),
    ("frame5", 512, 3, 3),
    ("frame6", 512, 1, 1),
    ("frame7", 512, 3, 4),
    ("frame8", 512, 1, 1),
    ("frame9", 512, 1, 1),
    ("frame10", 1500, 1, 1),
)


def xvector_param_shapes(input_dim, num_outputs, frame_layers=FRAME_LAYERS, output_name="outputs"):
    """Keras layouts: Conv1D kernel [k, C_in, C_out], Dense kernel [in, out] (xvector.py:38-65).
    frame_layers=XVECTOR_EXTENDED_FRAME_LAYERS, output_name="output" gives xvector_extended.py:22-43."""
    shapes = {}
    c_in = input_dim
    for name, filters, k, _ in frame_layers:
        shapes[name + "/kernel"] = (k, c_in, filters)
        shapes[name + "/bias"] = (filters,)
        c_in = filters
    shapes["segment1/kernel"] = (2 * c_in, 512)
    shapes["segment1/bias"] = (512,)
    shapes["segment2/kernel"] = (512, 512)
    shapes["segment2/bias"] = (512,)
    shapes[output_name + "/kernel"] = (512, num_outputs)
    shapes[output_name + "/bias"] = (num_outputs,)
    return shapes


def xvector_init(input_dim, num_outputs, seed=0, dtype=np.float32, bias_scale=0.0, frame_layers=FRAME_LAYERS,
                 output_name="outputs"):
    """Keras default init: glorot-uniform kernels, zero biases.
    bias_scale > 0 draws small random biases instead so that tests exercise the bias path."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in xvector_param_shapes(input_dim, num_outputs, frame_layers, output_name).items():
        if name.endswith("/kernel"):
            receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
            limit = np.sqrt(6.0 / (fan_in + fan_out))
            params[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
        else:
            params[name] = (bias_scale * rng.standard_normal(shape)).astype(dtype)
    return params


def conv1d_causal(x, kernel, bias, stride, relu=True):
    """Keras Conv1D(padding="causal") (xvector.py:38-39):
    left-pad k-1 zeros, VALID cross-correlation, T_out = ceil(T / stride)."""
    B, T, C = x.shape
    k, c_in, c_out = kernel.shape
    assert c_in == C
    xp = np.concatenate([np.zeros((B, k - 1, C), x.dtype), x], axis=1)
    t_out = -(-T // stride)
    idx = np.arange(t_out)[:, None] * stride + np.arange(k)[None, :]
    cols = xp[:, idx, :].reshape(B, t_out, k * C)
    y = cols @ kernel.reshape(k * C, c_out) + bias
    return np.maximum(y, 0) if relu else y


def stats_pooling(x):
    """GlobalMeanStddevPooling1D.call, xvector.py:30-35 — population variance, two-pass, clip 1e-10 before sqrt."""
    mean = x.mean(axis=1, keepdims=True)
    var = np.square(x - mean).mean(axis=1)
    std = np.sqrt(np.clip(var, x.dtype.type(STDDEV_SQRT_MIN_CLIP), np.finfo(x.dtype).max))
    return np.concatenate([mean[:, 0, :], std], axis=1)


def log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def xvector_forward(params, x, embedding=False, return_activations=False, frame_layers=FRAME_LAYERS,
                    output_name="outputs", output_activation="log_softmax"):
    """lidbox/models/xvector.py:46-67 (forward, no dropout) and :70-73 (embedding = pre-ReLU segment1);
    with the extended layer list: lidbox/models/xvector_extended.py:22-43."""
    dtype = x.dtype
    acts = {}
    h = x
    for name, _, _, stride in frame_layers:
        h = conv1d_causal(h, params[name + "/kernel"].astype(dtype), params[name + "/bias"].astype(dtype), stride)
        acts[name] = h
    h = stats_pooling(h)
    acts["stats_pooling"] = h
    h = h @ params["segment1/kernel"].astype(dtype) + params["segment1/bias"].astype(dtype)
    if embedding:
        return (h, acts) if return_activations else h
    h = np.maximum(h, 0)
    acts["segment1"] = h
    h = np.maximum(h @ params["segment2/kernel"].astype(dtype) + params["segment2/bias"].astype(dtype), 0)
    acts["segment2"] = h
    h = h @ params[output_name + "/kernel"].astype(dtype) + params[output_name + "/bias"].astype(dtype)
    acts[output_name] = h
    out = log_softmax(h) if output_activation == "log_softmax" else h
    return (out, acts) if return_activations else out


# --------------------------------------------------------------------------- #
# lidbox/losses.py
# --------------------------------------------------------------------------- #


def ap_theta(z, N):
    """SparseAngularProximity.theta, losses.py:42-49 — acos(z @ c_T) with c_T = one-hot rows => acos(z[:, :N])."""
    return np.arccos(np.asarray(z)[:, :N])


def ap_loss_per_sample(y, z, N, delta_weight=1.0):
    """SparseAngularProximity.call, losses.py:25-40 with rank-1 labels (the in-tree self-test's usage, :70,:97)."""
    z = np.asarray(z)
    y = np.asarray(y).reshape(-1)
    theta = ap_theta(z, N)
    theta_l = theta[np.arange(len(y)), y]
    deltas = theta_l[:, None] - theta
    sig = 1.0 / (1.0 + np.exp(-z.dtype.type(delta_weight) * deltas))
    mask = 1.0 - np.eye(N, dtype=z.dtype)[y]
    return (mask * sig).sum(axis=1)


def ap_loss(y, z, N, delta_weight=1.0):
    """Keras Loss.__call__ default reduction SUM_OVER_BATCH_SIZE = mean over the batch."""
    return ap_loss_per_sample(y, z, N, delta_weight).mean()


def sparse_xent_on_logprobs(y, logp):
    """CE used by the training config: -mean_b logp[b, y_b]."""
    y = np.asarray(y).reshape(-1)
    return -logp[np.arange(len(y)), y].mean()


# --------------------------------------------------------------------------- #
# torch-CPU twin (autograd) — used for gradient parity and as the timed CPU baseline
# --------------------------------------------------------------------------- #


def _bf16_ste(t, round_grad=False):
    """Round to bfloat16 in the forward pass with a straight-through gradient; optionally also round the incoming
    gradient to bfloat16 (mirrors where the bf16 training path stores activations / data gradients in bf16)."""
    import torch
    r = t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())
    if round_grad and r.requires_grad:
        r.register_hook(lambda g: g.to(torch.bfloat16).to(g.dtype))
    return r


def torch_xvector_forward(params, x, embedding=False, l2_normalize=False, emulate_bf16=False,
                          frame_layers=FRAME_LAYERS, output_name="outputs"):
    """Same math as xvector_forward with torch ops so autograd provides the backward.
    params: dict name -> torch tensor in Keras layouts; x: [B, T, F].
    emulate_bf16=True restates the precision="bf16" training path: weights, stored activations and stored data
    gradients are rounded to bfloat16 at the points where the CUDA path keeps them in bf16 (accumulation, biases,
    pooling statistics, logits and the loss stay in high precision).  It exists so that the backward kernels can be
    checked tightly; the plain fp64 path remains the reference semantics."""
    import torch
    import torch.nn.functional as F
    q = (lambda t, g=False: _bf16_ste(t, g)) if emulate_bf16 else (lambda t, g=False: t)
    h = q(x).transpose(1, 2)                                 # NCW for conv1d
    for name, _, k, stride in frame_layers:
        w = q(params[name + "/kernel"]).permute(2, 1, 0)     # [k, Cin, Cout] -> [Cout, Cin, k] (cross-correlation)
        h = q(F.relu(F.conv1d(F.pad(h, (k - 1, 0)), w, params[name + "/bias"], stride=stride)), True)
    mean = h.mean(dim=2)
    var = ((h - mean[:, :, None]) ** 2).mean(dim=2)
    std = torch.sqrt(torch.clamp(var, min=STDDEV_SQRT_MIN_CLIP))
    h = q(torch.cat([mean, std], dim=1))
    h = h @ q(params["segment1/kernel"]) + params["segment1/bias"]
    if embedding:
        return h
    h = q(F.relu(h), True)
    h = q(F.relu(h @ q(params["segment2/kernel"]) + params["segment2/bias"]), True)
    h = h @ q(params[output_name + "/kernel"]) + params[output_name + "/bias"]
    if emulate_bf16 and h.requires_grad:
        h.register_hook(lambda g: g.to(torch.bfloat16).to(g.dtype))
    if l2_normalize:                                         # spherespeaker.py:28-31 style head for the AP config
        return h / torch.sqrt(torch.clamp((h * h).sum(dim=1, keepdim=True), min=1e-12))
    return torch.log_softmax(h, dim=-1)


def torch_ap_loss(y, z, N, delta_weight=1.0):
    import torch
    theta = torch.acos(z[:, :N])
    theta_l = theta.gather(1, y.view(-1, 1).long())
    sig = torch.sigmoid(delta_weight * (theta_l - theta))
    mask = 1.0 - torch.nn.functional.one_hot(y.view(-1).long(), N).to(z.dtype)
    return (mask * sig).sum(dim=1).mean()


def torch_logmel(signals, sample_rate=16000, frame_length_ms=25, frame_step_ms=10, fft_length=512,
                 num_mel_bins=40, fmin=0.0, fmax=8000.0, power=2.0):
    """fp32 multi-threaded CPU log-mel (torch.fft.rfft on explicit frames; NOT torch.stft, whose framing
    differs from TF's).  Used only as the timed CPU baseline; values are checked against `logmel`."""
    import torch
    L = ms_to_frames(sample_rate, frame_length_ms)
    step = ms_to_frames(sample_rate, frame_step_ms)
    frames = signals.unfold(1, L, step) * torch.from_numpy(hann_window(L))
    S = torch.fft.rfft(frames, n=fft_length, dim=-1).abs() ** power
    W = torch.from_numpy(linear_to_mel_weight_matrix(num_mel_bins, fft_length // 2 + 1, sample_rate, fmin, fmax))
    return torch.log(S @ W + 1e-6)


# --------------------------------------------------------------------------- #
# lidbox/data/steps.py:579-632 create_signal_chunks, lidbox/util.py:41-57, lidbox/metrics.py
# --------------------------------------------------------------------------- #


def create_signal_chunks(signal, sample_rate, length_ms, step_ms, max_pad_ms=0):
    """steps.py:586-588, :600-615 for one signal: [N] -> [num_chunks, chunk_length] (float32 time arithmetic,
    int32 truncation, optional zero padding of the last chunk, tf.signal.frame with pad_end=False)."""
    signal = np.asarray(signal)
    sr = np.float32(sample_rate)
    chunk_length = int(np.int32(sr * np.float32(1e-3 * length_ms)))
    chunk_step = int(np.int32(sr * np.float32(1e-3 * step_ms)))
    max_pad = int(np.int32(sr * np.float32(1e-3 * max_pad_ms)))
    num_full_chunks = max(0, 1 + (signal.size - chunk_length) // chunk_step)      # python // floors like tf int32 //
    last_chunk_length = signal.size - num_full_chunks * chunk_step
    if last_chunk_length < chunk_length and chunk_length <= last_chunk_length + max_pad:
        signal = np.concatenate([signal, np.zeros(chunk_length - last_chunk_length, signal.dtype)])
    n = num_frames(signal.size, chunk_length, chunk_step)
    if n == 0:
        return np.zeros((0, chunk_length), signal.dtype)
    return np.stack([signal[c * chunk_step:c * chunk_step + chunk_length] for c in range(n)])


def merge_chunk_predictions(chunk_ids, predictions):
    """util.py:41-57 with the default stack_and_average: rows whose id shares everything before the last '-' are
    averaged; returns (sorted parent ids, [G, ...] means)."""
    groups = {}
    for cid, p in zip(chunk_ids, predictions):
        groups.setdefault(cid.rsplit('-', 1)[0], []).append(np.asarray(p))
    ids = sorted(groups)
    return ids, np.stack([np.stack(groups[i]).mean(axis=0) for i in ids])


class AverageDetectionCost:
    """lidbox/metrics.py:6-99 restated with numpy float32 counters."""

    def __init__(self, N, thresholds, C_miss=1.0, C_fa=1.0, P_tar=0.5):
        assert N >= 2, "C_avg is undefined for less than 2 classes."          # metrics.py:21
        self.thresholds = np.asarray(thresholds, np.float32)
        assert self.thresholds.ndim == 1                                      # metrics.py:22
        T = self.thresholds.size
        self.N, self.C_miss, self.C_fa, self.P_tar = N, C_miss, C_fa, P_tar
        self.fn, self.tp = np.zeros((N, T), np.float32), np.zeros((N, T), np.float32)
        self.fp_pairs, self.tn_pairs = np.zeros((N, N, T), np.float32), np.zeros((N, N, T), np.float32)

    def reset_states(self):
        for a in (self.fn, self.tp, self.fp_pairs, self.tn_pairs):
            a[...] = 0

    def update_state(self, true_positives, predictions):
        tpos = np.asarray(true_positives, np.float32)                         # metrics.py:56-61
        label = tpos.argmax(axis=-1)
        tpos = tpos[:, :, None]
        tneg = (~tpos.astype(bool)).astype(np.float32)
        pred = np.asarray(predictions, np.float32)[:, :, None]
        ppos = (pred >= self.thresholds).astype(np.float32)
        pneg = (pred < self.thresholds).astype(np.float32)
        self.tp += (ppos * tpos).sum(axis=0)                                  # metrics.py:63-66
        self.fn += (pneg * tpos).sum(axis=0)
        np.add.at(self.fp_pairs, label, ppos * tneg)                          # metrics.py:68-71 scatter_nd_add
        np.add.at(self.tn_pairs, label, pneg * tneg)

    def result_per_threshold(self):
        f32 = np.float32
        P_miss = _divide_no_nan(self.fn, self.fn + self.tp).mean(axis=0, dtype=f32)      # metrics.py:81-86
        P_fa = _divide_no_nan(_divide_no_nan(self.fp_pairs, self.fp_pairs + self.tn_pairs).sum(axis=1, dtype=f32),
                              f32(self.N - 1)).mean(axis=0, dtype=f32)                   # metrics.py:89-98
        return f32(self.C_miss * self.P_tar) * P_miss + f32(self.C_fa * (1 - self.P_tar)) * P_fa

    def result(self):
        return self.result_per_threshold().min()                              # metrics.py:105


class SparseAverageDetectionCost(AverageDetectionCost):
    """metrics.py:104-109: tf.one_hot(labels, N) (out-of-range labels give an all-zero row), then the dense update."""

    def update_state(self, true_positives, predictions):
        y = np.asarray(true_positives).astype(np.int64).reshape(-1)
        onehot = np.zeros((y.size, self.N), np.float32)
        ok = (y >= 0) & (y < self.N)
        onehot[np.arange(y.size)[ok], y[ok]] = 1
        super().update_state(onehot, predictions)


def cavg_by_definition(labels, scores, threshold, N, C_miss=1.0, C_fa=1.0, P_tar=0.5):
    """Independent statement of Li, Ma & Lee (2013) eq. 32 with python loops (small cases only): used to pin the
    counter-based implementation above.  Classes without any trial contribute 0, as divide_no_nan does."""
    labels, scores = np.asarray(labels), np.asarray(scores, np.float64)
    p_miss = 0.0
    p_fa = 0.0
    for l in range(N):
        target = scores[labels == l]                       # trials whose true class is l
        if len(target):
            p_miss += np.mean(target[:, l] < threshold)    # class l rejected although true
        fa_l = 0.0
        for m in range(N):
            if m == l or not len(target):
                continue
            fa_l += np.mean(target[:, m] >= threshold)     # class m accepted although the truth is l
        p_fa += fa_l / (N - 1)
    return C_miss * P_tar * p_miss / N + C_fa * (1 - P_tar) * p_fa / N
