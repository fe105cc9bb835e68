"""Drop-in for the signal-chunking step in front of the feature stage, lidbox/data/steps.py:579-632
(`create_signal_chunks`), as function-level building blocks (the tf.data orchestration around it is out of scope):

    chunk_geometry(sample_rate, length_ms, step_ms, max_pad_ms)      # steps.py:586-588, :603-605 (float32 arithmetic)
    signal_chunks(signals, sample_rate, length_ms, step_ms, max_pad_ms=0)   # [B, N] -> [B, C, L] on the device
    create_signal_chunks(elements, length_ms, step_ms, ...)          # dict elements in, chunk elements out

One kernel (lbx_signal_chunks_f32) writes all chunks of a batch; the zero padding of the last chunk (steps.py:609-611)
is produced by the same kernel, so no padded copy of the signal is made.
"""
import numpy as np
import torch

from .. import _lib


def chunk_geometry(sample_rate, length_ms, step_ms, max_pad_ms=0):
    """(chunk_length, chunk_step, max_pad) in samples, computed as the reference does: the millisecond arguments become
    float32 seconds, are multiplied by the float32 sample rate and truncated to int32."""
    sr = np.float32(sample_rate)
    length = int(np.int32(sr * np.float32(1e-3 * length_ms)))
    step = int(np.int32(sr * np.float32(1e-3 * step_ms)))
    pad = int(np.int32(sr * np.float32(1e-3 * max_pad_ms)))
    if length <= 0 or step <= 0:
        raise ValueError("chunk length and step must be at least one sample")
    return length, step, pad


def num_signal_chunks(num_samples, chunk_length, chunk_step, max_pad=0):
    n = _lib.lib().lbx_num_signal_chunks(int(num_samples), int(chunk_length), int(chunk_step), int(max_pad))
    if n < 0:
        raise ValueError("bad chunk geometry")
    return int(n)


def signal_chunks(signals, sample_rate, length_ms, step_ms, max_pad_ms=0):
    """signals [N] or [B, N] (equal lengths) -> chunks [C, L] or [B, C, L] (float32, device)."""
    sig = signals if isinstance(signals, torch.Tensor) else torch.as_tensor(np.asarray(signals))
    squeeze = sig.dim() == 1
    if squeeze:
        sig = sig[None]
    if sig.dim() != 2:
        raise ValueError("signals must be [N] or [B, N]")
    dev = _lib.require_cuda()
    sig = sig.to(dev, torch.float32).contiguous()
    B, N = sig.shape
    L, step, pad = chunk_geometry(sample_rate, length_ms, step_ms, max_pad_ms)
    C = num_signal_chunks(N, L, step, pad)
    out = torch.empty((B, C, L), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().lbx_signal_chunks_f32(_lib.ptr(sig), B, N, L, step, C, _lib.ptr(out), _lib.stream_ptr(dev)))
    return out[0] if squeeze else out


def chunk_ids(utterance_id, num_chunks, max_num_chunks_per_signal=int(1e6)):
    """steps.py:589-594: '<id>-<chunk number, 1-based, zero-filled to round(log10(max_num_chunks)) digits>'."""
    width = int(np.round(np.float32(np.log(np.float32(max_num_chunks_per_signal))) / np.float32(np.log(10.0))))
    return ["%s-%0*d" % (utterance_id, width, c) for c in range(1, num_chunks + 1)]


def create_signal_chunks(elements, length_ms, step_ms, max_pad_ms=0, max_num_chunks_per_signal=int(1e6)):
    """Generator over chunk elements: every element (dict with "id", "signal", "sample_rate", optionally "duration")
    is divided into fixed-length chunks; metadata is repeated, ids get the chunk number appended, "duration" is
    recomputed (steps.py:589-621)."""
    for x in elements:
        sig = x["signal"]
        n = int(sig.shape[-1]) if hasattr(sig, "shape") else len(sig)
        L, step, pad = chunk_geometry(x["sample_rate"], length_ms, step_ms, max_pad_ms)
        full = num_signal_chunks(n, L, step, 0)
        if full >= max_num_chunks_per_signal:
            raise ValueError("Too many chunks created from signal, cannot create unique utterance ids, raise the "
                             "max_num_chunks_per_signal parameter")
        chunks = signal_chunks(sig, x["sample_rate"], length_ms, step_ms, max_pad_ms)
        for cid, chunk in zip(chunk_ids(x["id"], chunks.shape[0], max_num_chunks_per_signal), chunks):
            out = dict(x, signal=chunk, id=cid)
            if "duration" in x:
                out["duration"] = float(np.float32(chunk.numel() / x["sample_rate"]))
            yield out
