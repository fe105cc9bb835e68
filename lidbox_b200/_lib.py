"""ctypes binding of liblidbox_b200.so (include/lidbox_b200.h).  There is no CPU fallback: if the library is
missing or a call fails, the error is raised to the caller."""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBX_LIB") or os.path.join(_PKG_DIR, "liblidbox_b200.so")   # LBX_LIB: A/B builds

c_int, c_ll, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_P = c_void_p
F32, BF16, I16 = 0, 1, 2      # dtype tags of include/lidbox_b200.h



class GemmDesc(ctypes.Structure):
    """lbx_gemm_t (include/lidbox_b200.h)."""
    _fields_ = [
        ("a0", c_void_p), ("a1", c_void_p), ("a_rows", c_ll), ("a_cols", c_int), ("lda", c_ll),
        ("b0", c_void_p), ("b1", c_void_p), ("b_rows", c_ll), ("b_cols", c_int), ("ldb", c_ll),
        ("layout", c_int), ("n_terms", c_int), ("term_a", c_int * 5), ("term_b", c_int * 5), ("term_a_row", c_int * 5),
        ("term_b_row", c_int * 5), ("term_col_limit", c_int * 5), ("b_map_rows", c_ll),
        ("k_splits", c_int), ("epi_atomic", c_int), ("out_dtype", c_int),
        ("out", c_void_p), ("out_lo", c_void_p), ("ldo", c_ll), ("bias", c_void_p), ("relu", c_int),
        ("post_scale", c_void_p), ("post_shift", c_void_p),
        ("rows_per_utt", c_int), ("valid_rows", c_int), ("mask_src", c_void_p), ("accumulate", c_int),
        ("tile_n", c_int), ("colsum", c_void_p), ("colsum_mod", c_int),
    ]


class WgradDesc(ctypes.Structure):
    """lbx_wgrad_t (include/lidbox_b200.h)."""
    _fields_ = [
        ("a", c_void_p), ("rows", c_ll), ("a_cols", c_int), ("lda", c_ll),
        ("b", c_void_p), ("b_cols", c_int), ("ldb", c_ll),
        ("out", c_void_p), ("ldo", c_ll),
    ]


class LogmelDesc(ctypes.Structure):
    """lbx_logmel_t (include/lidbox_b200.h)."""
    _fields_ = [
        ("sig", c_void_p), ("sig_dtype", c_int), ("B", c_ll), ("N", c_ll),
        ("frame_length", c_int), ("frame_step", c_int), ("fft_length", c_int), ("power", c_float),
        ("n_mel", c_int), ("band_start", c_void_p), ("band_len", c_void_p), ("band_off", c_void_p),
        ("band_w", c_void_p), ("n_packed", c_int), ("log_mode", c_int), ("eps", c_float),
        ("out", c_void_p), ("out_lo", c_void_p), ("out_dtype", c_int),
        ("out_utt_pitch", c_ll), ("out_row_pitch", c_int),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
    ]


# name -> (restype, argtypes); must list every symbol include/lidbox_b200.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "lbx_last_error": (ctypes.c_char_p, []),
    "lbx_version": (c_int, []),
    "lbx_launch_count": (c_ll, []),
    "lbx_ms_to_frames": (c_int, [c_int, c_int]),
    "lbx_num_frames": (c_ll, [c_ll, c_int, c_int]),
    "lbx_mel_weight_matrix": (c_int, [c_int, c_int, c_int, c_float, c_float, _P]),
    "lbx_mel_pack_bands": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "lbx_spectrogram_f32": (c_int, [_P, c_ll, c_ll, c_int, c_int, c_int, c_float, _P, _P]),
    "lbx_linear_to_mel_f32": (c_int, [_P, c_ll, c_int, c_int, _P, _P, _P, _P, c_int, c_int, c_float, _P, _P]),
    "lbx_logmel_workspace_bytes": (c_size_t, [c_ll, c_ll, c_int, c_int, c_int, c_int]),
    "lbx_logmel_f32": (c_int, [_P, c_ll, c_ll, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_int, c_int,
                               c_float, _P, _P, c_size_t, _P]),
    "lbx_logmel_i16": (c_int, [_P, c_ll, c_ll, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_int, c_int,
                               c_float, _P, _P]),
    "lbx_logmel_ex": (c_int, [ctypes.POINTER(LogmelDesc), _P]),
    "lbx_power_to_db_f32": (c_int, [_P, c_ll, c_float, c_float, _P, _P, _P]),
    "lbx_db_to_power_f32": (c_int, [_P, c_ll, _P, _P]),
    "lbx_check_finite_f32": (c_int, [_P, c_ll, _P, _P]),
    "lbx_normalize_axis_f32": (c_int, [_P, _P, c_ll, c_ll, c_ll, c_int, c_float, c_float, _P]),
    "lbx_feature_scaling_all_f32": (c_int, [_P, _P, c_ll, c_float, c_float, _P, _P]),
    "lbx_mfcc_f32": (c_int, [_P, c_ll, c_int, c_int, c_int, _P, _P]),
    "lbx_window_normalization_f32": (c_int, [_P, _P, c_ll, c_int, c_int, c_int, c_int, _P]),
    "lbx_row_rms_f32": (c_int, [_P, c_ll, c_int, _P, _P]),
    "lbx_rms_vad_f32": (c_int, [_P, c_ll, c_ll, c_int, c_float, c_float, c_ll, _P, _P, _P]),
    "lbx_vad_compact_f32": (c_int, [_P, c_ll, c_ll, c_int, _P, c_ll, _P, _P, _P, _P]),
    "lbx_num_signal_chunks": (c_ll, [c_ll, c_ll, c_ll, c_ll]),
    "lbx_signal_chunks_f32": (c_int, [_P, c_ll, c_ll, c_ll, c_ll, c_ll, _P, _P]),
    "lbx_group_mean_f32": (c_int, [_P, _P, _P, c_ll, c_int, _P, _P]),
    "lbx_cavg_update_f32": (c_int, [_P, _P, _P, c_ll, c_int, _P, c_int, _P, _P, _P, _P, _P]),
    "lbx_cavg_result_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, c_float, c_float, _P, _P, _P]),
    "lbx_gemm_bf16": (c_int, [ctypes.POINTER(GemmDesc), _P]),
    "lbx_set_pdl": (c_int, [c_int]),
    "lbx_dense_xent_head": (c_int, [_P, _P, _P, _P, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P,
                                    _P, _P]),
    "lbx_set_gemm_pair": (c_int, [c_int]),
    "lbx_set_gemm_fast_epilogue": (c_int, [c_int]),
    "lbx_pack_rows_bf16": (c_int, [_P, c_ll, c_int, c_int, _P, _P, c_int, c_int, c_int, c_float, ctypes.c_ulonglong, _P,
                                   _P]),
    "lbx_counter_tick": (c_int, [_P, _P]),
    "lbx_stats_pool_fwd": (c_int, [_P, c_int, c_ll, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "lbx_stats_pool_bwd": (c_int, [_P, c_ll, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P, c_int, _P]),
    "lbx_logsoftmax_xent": (c_int, [_P, _P, c_ll, c_int, _P, _P, _P, c_int, c_float, _P, _P]),
    "lbx_ap_loss": (c_int, [_P, _P, c_ll, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P, c_int, _P, c_float, _P, _P]),
    "lbx_wgrad_grouped": (c_int, [_P, c_int, _P]),
    "lbx_set_gemm_wide_epilogue": (c_int, [c_int, c_int]),
    "lbx_set_wgrad_quad": (c_int, [c_int]),
    "lbx_wgrad_grouped_plan": (c_int, [_P, c_int, c_int, c_int, _P, c_int, _P]),
    "lbx_head_fwd": (c_int, [_P, c_ll, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, _P, _P, _P, c_ll, _P, _P]),
    "lbx_head_bwd": (c_int, [_P, _P, _P, c_ll, c_int, c_int, c_int, _P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "lbx_adam_step": (c_int, [_P, _P, _P, _P, c_ll, c_float, c_float, c_float, c_float, _P, _P, c_float, _P, c_int, _P]),
    "lbx_adam_step_sharded": (c_int, [_P, _P, _P, _P, _P, _P, c_ll, c_int, c_int, _P, _P, c_float, c_float, c_float,
                                      c_float, _P, _P, c_float, c_int, _P, _P, _P, c_ll, _P]),
    "lbx_dp_signal": (c_int, [_P, c_int, c_int, c_int, _P, ctypes.c_uint, _P]),
    "lbx_dp_wait_slot": (c_int, [_P, c_int, c_int, _P, ctypes.c_uint, _P, _P]),
    "lbx_dp_wait": (c_int, [_P, c_int, _P, _P, _P]),
    "lbx_set_dp_spin_limit": (c_int, [c_ll]),
    "lbx_set_dp_blocks_per_sm": (c_int, [c_int]),
    "lbx_split_bf16": (c_int, [_P, c_ll, _P, _P, _P]),
    "lbx_logmel_f32_host": (c_int, [_P, c_ll, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_float,
                                    c_int, c_float, _P, _P, _P, _P, c_size_t, _P]),
}

_lib = None


class LidboxB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LidboxB200Error(
                "%s is missing: build it with `python -m lidbox_b200.build` (nvcc, sm_100a). "
                "lidbox_b200 has no CPU or library fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if os.environ.get("LBX_PDL", "1") == "0":
            handle.lbx_set_pdl(0)
        handle.lbx_set_gemm_pair(0 if os.environ.get("LBX_GEMM_PAIR", "1") == "0" else 1)
        handle.lbx_set_gemm_fast_epilogue(0 if os.environ.get("LBX_GEMM_FAST_EPI", "1") == "0" else 1)
        handle.lbx_set_wgrad_quad(1 if os.environ.get("LBX_WGRAD_QUAD", "0") == "1" else 0)
        handle.lbx_set_gemm_wide_epilogue(0 if os.environ.get("LBX_GEMM_WIDE_EPI", "1") == "0" else 1,
                                          int(os.environ.get("LBX_GEMM_WIDE_MAX_KB", "0")))
        if os.environ.get("LBX_DP_BLOCKS_PER_SM"):
            handle.lbx_set_dp_blocks_per_sm(int(os.environ["LBX_DP_BLOCKS_PER_SM"]))
        if os.environ.get("LBX_DP_SPIN_LIMIT"):
            handle.lbx_set_dp_spin_limit(int(os.environ["LBX_DP_SPIN_LIMIT"]))
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().lbx_last_error()
        text = msg.decode() if msg else ""
        if rc == -2:
            raise NotImplementedError("lidbox_b200: " + text)
        if rc == -1:
            raise ValueError("lidbox_b200: " + text)
        raise LidboxB200Error("lidbox_b200 error %d: %s" % (rc, text))


def ptr(t):
    """Device (or host) pointer of a contiguous tensor as a void*; None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous()
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise LidboxB200Error("lidbox_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    cur = torch.device("cuda", torch.cuda.current_device())
    if device is None:
        return cur
    dev = torch.device(device)
    if dev.type != "cuda":
        raise LidboxB200Error("lidbox_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if dev.index is not None and dev.index != cur.index:
        # the library launches on the CURRENT device and keeps its per-kernel attributes / SM counts per process (one
        # process per GPU): a model on another ordinal would launch on the wrong GPU
        raise LidboxB200Error("device %s is not the current CUDA device (%s): call torch.cuda.set_device() first "
                              "(one process per GPU)" % (dev, cur))
    return cur
