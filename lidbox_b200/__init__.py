"""lidbox_b200 — B200-native (sm_100a) drop-in for the lidbox log-mel + x-vector hot path.

Mirrors the call surface of lidbox.features.audio, lidbox.features.mel_ops, lidbox.models.xvector, lidbox.losses and
lidbox.data.tf_utils.extract_features with torch CUDA tensors in place of tf.Tensor.  All arithmetic runs in
hand-written CUDA kernels behind the C-ABI in include/lidbox_b200.h; there is no CPU or library fallback.
"""
__version__ = "0.1.0"
