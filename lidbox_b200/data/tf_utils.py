"""Drop-in for the map stage lidbox/data/tf_utils.py:166-195 (`extract_features`), the call site
`steps.extract_features` (lidbox/data/steps.py:708-736) invokes once per batch.

Same positional signature and error convention (rank / sample-rate / finiteness checks raise).  The reference calls
`audio_features.melspectrograms`, a name that does not exist at this commit; the intended chain
spectrograms -> linear_to_mel -> log(x + 1e-6) is implemented, fused into one kernel for the (log-)mel feature types.
Re-entrant: every call works on the calling thread's current CUDA stream and shares only read-only cached tables.
"""
import numpy as np
import torch

from .. import features
from ..features import audio as audio_features


def extract_features(signals, sample_rates, feattype, spec_kwargs=None, melspec_kwargs=None, mfcc_kwargs=None,
                     db_spec_kwargs=None, feat_scale_kwargs=None, window_norm_kwargs=None, check_finite=True):
    spec_kwargs, melspec_kwargs = dict(spec_kwargs or {}), dict(melspec_kwargs or {})
    sig = signals if isinstance(signals, torch.Tensor) else torch.as_tensor(np.asarray(signals))
    if sig.dim() != 2:
        raise ValueError("Input signals for feature extraction must be batches of mono signals without channels, "
                         "i.e. of shape [B, N] where B is batch size and N number of samples.")
    rates = np.asarray(sample_rates.cpu() if isinstance(sample_rates, torch.Tensor) else sample_rates).reshape(-1)
    if rates.size == 0 or not (rates == rates[0]).all():
        raise ValueError("Different sample rates in a single batch not supported, all signals in the same batch "
                         "should have the same sample rate.")
    sample_rate = int(rates[0])
    if feattype in ("melspectrogram", "logmelspectrogram"):
        X = audio_features.logmelspectrograms(sig, sample_rate, log=(feattype == "logmelspectrogram"), **spec_kwargs,
                                              **melspec_kwargs)
    elif feattype == "db_spectrogram":
        X = audio_features.spectrograms(sig, sample_rate, **spec_kwargs)
        if check_finite:
            audio_features.assert_all_finite(X, "spectrogram failed")
        X = audio_features.power_to_db(X, **(db_spec_kwargs or {}))
    elif feattype == "mfcc":
        X = audio_features.logmelspectrograms(sig, sample_rate, log=True, **spec_kwargs, **melspec_kwargs)
        if check_finite:
            audio_features.assert_all_finite(X, "logmelspectrogram failed")
        mfcc_kwargs = mfcc_kwargs or {}
        X = audio_features.mfccs_from_log_mel_spectrograms(X, mfcc_kwargs.get("coef_begin", 1),
                                                           mfcc_kwargs.get("coef_end", 13))   # tf_utils.py:181-185
    else:
        # "spectrogram" and, exactly like the reference's if/elif chain (tf_utils.py:172-188), ANY other feature type
        # string: the power spectrogram is returned unchanged, no error is raised
        X = audio_features.spectrograms(sig, sample_rate, **spec_kwargs)
    if check_finite:
        audio_features.assert_all_finite(X, str(feattype) + " failed")
    if feat_scale_kwargs:
        X = features.feature_scaling(X, **feat_scale_kwargs)                      # tf_utils.py:189-191
        if check_finite:
            audio_features.assert_all_finite(X, "feature scaling failed")
    if window_norm_kwargs:
        X = features.window_normalization(X, **window_norm_kwargs)                 # tf_utils.py:192-194
        if check_finite:
            audio_features.assert_all_finite(X, "window normalization failed")
    return X
