/*
 * lidbox_b200 — C-ABI of the B200-native (sm_100a) log-mel + x-vector hot path.
 *
 * The reference (py-lidbox/lidbox @ e60d5ad) has no FFI of its own: its operator API for this path is the
 * Python call surface of lidbox/features/audio.py, lidbox/features/mel_ops.py, lidbox/models/xvector.py,
 * lidbox/losses.py and lidbox/data/tf_utils.py, all dispatching into TensorFlow.  Each entry point below
 * replaces the TensorFlow op chain behind one of those call sites (cited as file:line relative to
 * /root/reference); the Python host package `lidbox_b200` binds them with ctypes.
 *
 * Conventions
 *  - every function returns int: 0 = OK, <0 = LBX_E* code; lbx_last_error() gives a thread-local message;
 *  - all tensor pointers are DEVICE pointers owned by the caller unless the name ends in `_host`;
 *    the library never allocates, frees or synchronises the device (the *_host convenience entry points are the
 *    only exception: they stage through caller-visible pinned/pageable host buffers and synchronise the stream);
 *  - all work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy default stream);
 *  - tensors are contiguous row-major; `long long` is used for sizes;
 *  - re-entrant and thread-safe: no mutable global state except write-once function attributes.
 */
#ifndef LIDBOX_B200_H
#define LIDBOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBX_OK 0
#define LBX_EINVAL (-1)      /* bad argument (shape, dtype, alignment) */
#define LBX_EUNSUPPORTED (-2) /* valid in the reference but not implemented by this library */
#define LBX_ECUDA (-3)       /* CUDA runtime / driver error */
#define LBX_EWORKSPACE (-4)  /* workspace missing or too small */

/* dtype tags used by the TDNN entry points */
#define LBX_F32 0
#define LBX_BF16 1
#define LBX_I16 2            /* 16-bit PCM samples (feature entry points only) */

const char* lbx_last_error(void);
int lbx_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches claim) */
long long lbx_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Host-side integer / table helpers
 * ---------------------------------------------------------------------------------------------------------- */

/* lidbox/features/audio.py:185-189  ms_to_frames: int32(float32(sr) * 1e-3f * float32(ms)) */
int lbx_ms_to_frames(int sample_rate, int ms);
/* tf.signal.frame(pad_end=False) as used by tf.signal.stft at audio.py:229: max(0, 1 + (N - L) / step) */
long long lbx_num_frames(long long n_samples, int frame_length, int frame_step);
/* lidbox/features/mel_ops.py:28-75 (bug-compatible, fp32): writes W_host[n_bins * n_mel] row-major [n_bins, n_mel] */
int lbx_mel_weight_matrix(int n_mel, int n_bins, int sample_rate, float lower_edge_hertz, float upper_edge_hertz,
                          float* W_host);
/* Band-compress a [n_bins, n_mel] weight matrix: for every mel column the contiguous range of non-zero rows.
 * start_host/len_host/off_host have n_mel entries, packed_host must hold n_bins*n_mel floats in the worst case;
 * returns the number of packed weights (>= 0) or <0 on error. */
int lbx_mel_pack_bands(const float* W_host, int n_bins, int n_mel, int* start_host, int* len_host, int* off_host,
                       float* packed_host);

/* ------------------------------------------------------------------------------------------------------------
 * Feature kernels (replace tf.signal.stft / abs / pow / tensordot / log behind lidbox.features.audio)
 * ---------------------------------------------------------------------------------------------------------- */

/* lidbox/features/audio.py:219-230  spectrograms(): frame + periodic Hann + rFFT(fft_length) + |.|^power.
 * sig [B, N] f32 -> out [B, T, fft_length/2+1] f32, T = lbx_num_frames(N, frame_length, frame_step).
 * fft_length must be a power of two in [32, 4096] and >= frame_length. */
int lbx_spectrogram_f32(const float* sig, long long B, long long N, int frame_length, int frame_step, int fft_length,
                        float power, float* out, void* stream);

/* lidbox/features/audio.py:247-261  linear_to_mel(): S [rows, n_bins] x band-packed W -> out [rows, n_mel].
 * log_mode 0: none; 1: ln(x + eps) (lidbox/data/tf_utils.py:178). */
int lbx_linear_to_mel_f32(const float* S, long long rows, int n_bins, int n_mel, const int* band_start,
                          const int* band_len, const int* band_off, const float* band_w, int n_packed, int log_mode,
                          float eps, float* out, void* stream);

/* Fused spectrograms -> linear_to_mel -> [ln(x+eps)] (tf_utils.py:172-178) in one pass: sig [B,N] -> out [B,T,n_mel].
 * Same argument rules as lbx_spectrogram_f32. workspace is needed only when the fused 512-point fast path does not
 * apply (lbx_logmel_workspace_bytes() > 0); it then holds the intermediate power spectrogram. */
size_t lbx_logmel_workspace_bytes(long long B, long long N, int frame_length, int frame_step, int fft_length,
                                  int n_mel);
int lbx_logmel_f32(const float* sig, long long B, long long N, int frame_length, int frame_step, int fft_length,
                   float power, int n_mel, const int* band_start, const int* band_len, const int* band_off,
                   const float* band_w, int n_packed, int log_mode, float eps, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Same fused chain fed with 16-bit PCM (the sample format of the WAV corpora; lidbox/features/audio.py:17-33 read_wav
 * decodes it to float32 as x / 32768, which is what the kernel does while staging): halves the host->device bytes of
 * the input pipeline.  Fused 512-point configuration only. */
int lbx_logmel_i16(const short* pcm, long long B, long long N, int frame_length, int frame_step, int fft_length,
                   float power, int n_mel, const int* band_start, const int* band_len, const int* band_off,
                   const float* band_w, int n_packed, int log_mode, float eps, float* out, void* stream);

/* The same fused chain with every option in one descriptor — the entry point the input pipeline binds
 * (lidbox/data/tf_utils.py:172-178 followed directly by the first Conv1D of lidbox/models/xvector.py:53):
 *   sig_dtype LBX_F32 | LBX_I16 (PCM decoded as x / 32768, audio.py:17-33);
 *   out_dtype LBX_F32 | LBX_BF16; frame t of utterance b is written at out + b*out_utt_pitch + t*out_row_pitch
 *   (elements; 0 = dense [B,T,n_mel]), so bf16 rows can land straight in the zero-left-padded activation buffer of
 *   the first frame layer (no fp32 [B,T,n_mel] round trip, no packing pass); out_lo (optional) receives the bf16
 *   residual x - bf16(x) for the bf16x3 forward mode.  Columns n_mel..out_row_pitch-1 are not written.
 * workspace: as for lbx_logmel_f32 (only when the fused 512-point path does not apply; then f32 dense output only). */
typedef struct lbx_logmel_t {
  const void* sig; int sig_dtype;
  long long B; long long N;
  int frame_length; int frame_step; int fft_length; float power;
  int n_mel; const int* band_start; const int* band_len; const int* band_off; const float* band_w; int n_packed;
  int log_mode; float eps;
  void* out; void* out_lo; int out_dtype;
  long long out_utt_pitch; int out_row_pitch;
  void* workspace; size_t workspace_bytes;
} lbx_logmel_t;
int lbx_logmel_ex(const lbx_logmel_t* d, void* stream);

/* lidbox/features/audio.py:167-174  power_to_db(): 20*(log10(max(amin,S)) - log10(max(amin,max_all S))),
 * floored at max_all(db) - top_db.  workspace: >= 16 bytes of device memory. */
int lbx_power_to_db_f32(const float* S, long long numel, float amin, float top_db, float* out, void* workspace,
                        void* stream);

/* lidbox/features/audio.py:177-181  db_to_power(): pow(10, S / 20), elementwise */
int lbx_db_to_power_f32(const float* S, long long numel, float* out, void* stream);

/* tf.debugging.assert_all_finite (tf_utils.py:173-194): writes 1 to *flag_dev (int32, device) if any element is
 * NaN/Inf, leaves it untouched otherwise (caller zeroes it first). */
int lbx_check_finite_f32(const float* x, long long numel, int* flag_dev, void* stream);

/* Host-buffer entry point (the e2e path a non-torch caller binds): pageable or pinned HOST signals in,
 * HOST log-mel out; device staging buffers are supplied by the caller (dev_sig >= B*N floats, dev_out >= B*T*n_mel). */
int lbx_logmel_f32_host(const float* sig_host, long long B, long long N, int sample_rate, int frame_length_ms,
                        int frame_step_ms, int fft_length, float power, int n_mel, float fmin, float fmax,
                        int log_mode, float eps, float* out_host, float* dev_sig, float* dev_out, void* dev_tables,
                        size_t dev_tables_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Feature normalisation (the step after log-mel in every pipeline: tf_utils.py:189-194, steps.py:821-834)
 * ---------------------------------------------------------------------------------------------------------- */

/* lidbox/features/__init__.py:16-20 cmn (mode 0), :26-32 cmvn (mode 1), :5-9 feature_scaling over one axis (mode 2,
 * result in [lo, hi]).  The tensor is viewed as [outer, R, inner] with R the reduced axis; fp32, two-pass statistics,
 * population std, divide_no_nan semantics. */
int lbx_normalize_axis_f32(const float* x, float* y, long long outer, long long R, long long inner, int mode, float lo,
                           float hi, void* stream);
/* feature_scaling with axis=None (min/max over the whole tensor); workspace: >= 8 bytes of device memory */
int lbx_feature_scaling_all_f32(const float* x, float* y, long long n, float lo, float hi, void* workspace,
                                void* stream);
/* lidbox/data/tf_utils.py:180-185: tf.signal.mfccs_from_log_mel_spectrograms(X)[..., coef_begin:coef_end] —
 * orthonormal-scaled DCT-II of the log-mel rows: out [rows, coef_end-coef_begin] */
int lbx_mfcc_f32(const float* logmel, long long rows, int n_mel, int coef_begin, int coef_end, float* out, void* stream);
/* lidbox/features/__init__.py:40-67 window_normalization over the time axis of [B,T,F] for 1 <= window_len < T
 * (REFLECT padding w/2 left, w/2-1+(w&1) right; mean / population std over each window) */
int lbx_window_normalization_f32(const float* x, float* y, long long B, int T, int F, int window_len,
                                 int normalize_variance, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Chunking, chunk-score merging and C_avg (the callers either side of the path: SURVEY.md §8(f) rows 2-4)
 * ---------------------------------------------------------------------------------------------------------- */

/* lidbox/data/steps.py:607-615 create_signal_chunks: number of chunks of chunk_length samples every chunk_step samples,
 * including the zero-padded last one when its missing tail is <= max_pad samples (host arithmetic; <0 on bad input) */
long long lbx_num_signal_chunks(long long N, long long chunk_length, long long chunk_step, long long max_pad);
/* steps.py:612-615: sig [B,N] -> out [B*num_chunks, chunk_length]; chunk c of signal b is row b*num_chunks + c, samples
 * behind the end of the signal are zero */
int lbx_signal_chunks_f32(const float* sig, long long B, long long N, long long chunk_length, long long chunk_step,
                          long long num_chunks, float* out, void* stream);
/* lidbox/util.py:41-57 merge_chunk_predictions (stack_and_average): out[g,:] = mean of the rows
 * row_index[group_offsets[g] .. group_offsets[g+1]) of pred [rows, D] */
int lbx_group_mean_f32(const float* pred, const long long* row_index, const long long* group_offsets,
                       long long num_groups, int D, float* out, void* stream);
/* lidbox/metrics.py:52-72 AverageDetectionCost.update_state (onehot [B,N] f32, labels NULL) and
 * :104-109 SparseAverageDetectionCost.update_state (labels [B] i32, onehot NULL): adds the batch to the counters
 * tp, fn [N,Th] and fp_pairs, tn_pairs [N,N,Th] (f32, caller-zeroed at reset_states) for scores pred [B,N]. */
int lbx_cavg_update_f32(const float* onehot, const int* labels, const float* pred, long long B, int N,
                        const float* thresholds, int num_thresholds, float* tp, float* fn, float* fp_pairs,
                        float* tn_pairs, void* stream);
/* metrics.py:74-99 result(): C_avg per threshold (optional output, [Th]) and its minimum (device scalar) */
int lbx_cavg_result_f32(const float* tp, const float* fn, const float* fp_pairs, const float* tn_pairs, int N,
                        int num_thresholds, float C_miss, float C_fa, float P_tar, float* cavg_per_threshold,
                        float* cavg_min, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Energy VAD (the step before the feature stage: steps.py:417-432 compute_rms_vad, :183-200 apply_vad)
 * ---------------------------------------------------------------------------------------------------------- */

/* lidbox/features/audio.py:264-271 root_mean_square over the last axis of a [rows, len] matrix */
int lbx_row_rms_f32(const float* x, long long rows, int len, float* out, void* stream);
/* audio.py:307-329 framewise_rms_energy_vad_decisions, batched: sig [B,N] -> decisions [B,F] (1 = voiced), F = N /
 * frame_step non-overlapping frames; threshold = strength * max(min_rms_threshold, mean RMS of the utterance); runs of
 * non-speech shorter than min_non_speech_frames are flipped back to speech (audio.py:289-296). rms_ws: [B,F] floats. */
int lbx_rms_vad_f32(const float* sig, long long B, long long N, int frame_step, float strength,
                    float min_rms_threshold, long long min_non_speech_frames, unsigned char* decisions, float* rms_ws,
                    void* stream);
/* audio.py:337-353 remove_silence / steps.py:191-198: keep the voiced frames of every utterance, compacted to the front
 * of out [B,N]; out_len [B] = voiced samples; offsets_ws: [B,F] int64. */
int lbx_vad_compact_f32(const float* sig, long long B, long long N, int frame_len, const unsigned char* decisions,
                        long long F, float* out, long long* out_len, long long* offsets_ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * TDNN contractions (replace Keras Conv1D / Dense behind lidbox/models/xvector.py:38-43,53-64 and their gradients)
 * ---------------------------------------------------------------------------------------------------------- */

/* One bf16 tensor-core GEMM (tcgen05, fp32 accumulation in TMEM) with a fused epilogue.
 *   layout 0 (NT):  C[M,N] = A[M,K] . B[N,K]^T   M = a_rows, K = a_cols = b_cols, N = b_rows
 *   layout 1 (TN):  C[M,N] = A[K,M]^T . B[K,N]   K = a_rows = b_rows, M = a_cols, N = b_cols
 *   layout 2 (NN):  C[M,N] = A[M,K] . B[K,N]     M = a_rows, K = a_cols = b_rows, N = b_cols (forward pass straight from
 *                                                the Keras-layout weight [k*C_in, C_out]: no transposed copy is kept)
 * A/B are bf16 views: `rows` x `cols` with row pitch ld (elements, multiple of 8; may be SMALLER than cols: a causal
 * Conv1D with kernel k and stride s over NWC activations is the NT GEMM whose A view has cols = k*C_in and
 * lda = s*C_in on the zero-left-padded activation buffer — no im2col).
 * n_terms (1..LBX_GEMM_MAX_TERMS) accumulating passes over the contraction; pass t reads A plane term_a[t] (0: a0,
 * 1: a1) shifted by term_a_row[t] rows, and B plane term_b[t] (0: b0, 1: b1) shifted by term_b_row[t] rows of the B view
 * (whose full extent is then b_map_rows >= term_b_row[t] + b_rows).  Three uses:
 *   dilated causal Conv1D (dilation_rate d, stride 1): tap j is the pass that reads the activations j*d rows further
 *   down against rows [j*C_in, (j+1)*C_in) of the Keras kernel (and the mirror image in its data gradient);
 *   "bf16x3": a1/b1 = bf16 residual planes (x - bf16(x)), passes (a0,b0) (a0,b1) (a1,b0): fp32-grade forward results;
 *   gather-form data gradient of a strided conv (k > stride): output time tau = t*stride + j receives tap j of row t,
 *   so pass i reads dZ shifted by -i rows against the kernel rows of taps [i*stride, (i+1)*stride) — ONE GEMM with
 *   N = stride*C_in and ceil(k/stride) passes; kernel rows past k*C_in (b_map_rows) read as zeros (no read-modify-write
 *   pass, no second launch).
 * Epilogue, per output element (m, n), in this order: + bias[n]; ReLU; zero unless mask_src[m*ldo+n] > 0;
 * then either atomicAdd into fp32 out (epi_atomic, required for k_splits > 1), or out (+)= x as fp32 / bf16
 * (plus out_lo = bf16 residual).  Rows with (m % rows_per_utt) >= valid_rows are stored as ZERO when rows_per_utt > 0 (they land on
 * destination rows that must stay zero: junk rows and the next utterance's causal padding). */
#define LBX_GEMM_MAX_TERMS 5
typedef struct lbx_gemm_t {
  const void* a0; const void* a1;
  long long a_rows; int a_cols; long long lda;
  const void* b0; const void* b1;
  long long b_rows; int b_cols; long long ldb;
  int layout;
  int n_terms;
  int term_a[LBX_GEMM_MAX_TERMS]; int term_b[LBX_GEMM_MAX_TERMS]; int term_a_row[LBX_GEMM_MAX_TERMS];
  int term_b_row[LBX_GEMM_MAX_TERMS];
  int term_col_limit[LBX_GEMM_MAX_TERMS];   /* > 0: pass t only contributes to output columns < limit (multiple of 256): */
                                            /* tiles beyond it skip the pass instead of multiplying by zero-filled rows  */
  long long b_map_rows;     /* 0 = b_rows */
  int k_splits;
  int epi_atomic;
  int out_dtype;            /* LBX_F32 | LBX_BF16 */
  void* out; void* out_lo; long long ldo;
  const float* bias;
  int relu;
  const float* post_scale;  /* optional [N]: x = x * post_scale[n] + post_shift[n] AFTER the activation (inference-time */
  const float* post_shift;  /* BatchNormalization behind a frame layer: conv -> ReLU -> BN, lidbox/models/xvector_2d.py:41-43) */
  int rows_per_utt; int valid_rows;
  const void* mask_src;     /* bf16, indexed like out */
  int accumulate;
  int tile_n;               /* 0 = automatic, or 64 / 128 / 256 */
  float* colsum;            /* optional: colsum[n % colsum_mod] += sum_m x[m,n] of the masked result (bias gradient) */
  int colsum_mod;
} lbx_gemm_t;
int lbx_gemm_bf16(const lbx_gemm_t* g, void* stream);
/* Grouped weight gradients: out_p[a_cols, b_cols] += A_p^T . B_p for up to 8 problems in ONE persistent launch
 * (replaces the Keras-computed kernel gradients of the Conv1D frame layers, lidbox/models/xvector.py:38-43 under
 * keras_utils.py:135-147 `fit`).  A_p = layer input [rows, a_cols] bf16 through its (possibly overlapping-row) view of
 * pitch lda, B_p = gradient w.r.t. the layer's pre-activation [rows, b_cols] bf16 (pitch ldb), out_p fp32 with pitch
 * ldo, accumulated with vector atomics (the caller zeroes it once per step).  The k-blocks of all problems are cut into
 * equal contiguous ranges, one per CTA pair (stream-K), so no problem pays its own pipeline fill / tail. */
typedef struct {
  const void* a; long long rows; int a_cols; long long lda;
  const void* b; int b_cols; long long ldb;
  float* out; long long ldo;
} lbx_wgrad_t;
int lbx_wgrad_grouped(const lbx_wgrad_t* problems, int n, void* stream);
/* Host-side replay of lbx_wgrad_grouped's work partition (no launch, no GPU needed; the same decode function the kernel
 * runs): the k-blocks of all problems cut into `workers` contiguous ranges.  Fills `segments` with rows of 6 ints
 * (worker, problem, m_unit, n_tile, first k-block, end k-block); pointers in `problems` are not dereferenced. */
int lbx_wgrad_grouped_plan(const lbx_wgrad_t* problems, int n, int workers, int quad, int* segments, int max_segments,
                           int* n_segments);
/* 1: clusters of two CTA pairs that share the A tile through TMA multicast (every problem needs an even number of
 * 256-column tiles); 0 (default): independent CTA pairs */
int lbx_set_wgrad_quad(int enabled);
/* GEMMs are launched with programmatic dependent launch (prologue overlaps the previous kernel's tail); 0 disables. */
int lbx_set_pdl(int enabled);
/* 1: the 256-wide GEMM tiles run on CTA pairs (clusters of 2, tcgen05 cta_group::2: a 256x256 tile per pair, every
 * CTA stages half of the B operand); 0: single-CTA 128x256 tiles. */
int lbx_set_gemm_pair(int enabled);
/* 1 (default): lean-epilogue CTA-pair launches with at most max_kb (default 64) k-blocks per tile run with 16 epilogue
 * warps instead of 8 (their pace is set by draining the accumulators, not by the tensor pipe); 0 disables */
int lbx_set_gemm_wide_epilogue(int enabled, int max_kb);
/* 1 (default): bf16-output GEMMs use the lean epilogue (bulk tensor stores / mask loads); 0: the general epilogue */
int lbx_set_gemm_fast_epilogue(int enabled);

/* ------------------------------------------------------------------------------------------------------------
 * TDNN non-GEMM stages and losses
 * ---------------------------------------------------------------------------------------------------------- */

/* features [B,T,F] f32 -> bf16 activation rows of the first frame layer: element (b,t,c) goes to row
 * b*rows_per_utt + row_off + t, column c of a [*, pitch] buffer (hi, and lo = residual when lo != NULL);
 * columns F..pitch-1 are written as zero.  drop_rate > 0 applies SpatialDropout1D (xvector.py:50-51): whole
 * channels of a sample are zeroed with probability drop_rate, the rest scaled by 1/(1-drop_rate).
 * The mask of a call is a hash of (seed + 7919 * *seed_counter_dev, sample, channel); seed_counter_dev (optional) is a
 * device-side call counter advanced by lbx_counter_tick(), so that replays of a captured CUDA graph draw a new mask. */
int lbx_pack_rows_bf16(const float* x, long long B, int T, int F, void* hi, void* lo, int rows_per_utt, int row_off,
                       int pitch, float drop_rate, unsigned long long seed, const unsigned long long* seed_counter_dev,
                       void* stream);
/* *counter_dev += 1, enqueued on the stream (graph-capturable) */
int lbx_counter_tick(unsigned long long* counter_dev, void* stream);

/* lidbox/models/xvector.py:25-35 GlobalMeanStddevPooling1D: y [B*rows_per_utt, pitch] (first T rows of every
 * utterance, C channels; f32 or bf16) -> out [B, 2C] f32 (mean | std), two-pass population variance in fp32,
 * std = sqrt(clip(var, clip_min, FLT_MAX)).  Optional: var_raw [B,C] (needed by the backward), bf16 copies. */
int lbx_stats_pool_fwd(const void* y, int y_dtype, long long B, int rows_per_utt, int T, int C, int pitch,
                       float clip_min, float* out, float* var_raw, void* out_hi, void* out_lo, void* stream);
/* backward of the pooling fused with the ReLU mask of the producing layer: dz = (y > 0) * d pool / d y . gpool */
/* dbias (optional): bias gradient of the producing layer, dbias[c] += sum_{b,t} dz[b,t,c]; zero_gpool: reset gpool
 * after it has been consumed (it is accumulated atomically by the split-K GEMM of the next step). */
int lbx_stats_pool_bwd(const void* y_bf16, long long B, int rows_per_utt, int T, int C, int pitch, float clip_min,
                       const float* pooled, const float* var_raw, float* gpool, void* dz_bf16, float* dbias,
                       int zero_gpool, void* stream);

/* log_softmax (xvector.py:64-65) + sparse categorical cross-entropy on the log-probs, forward + gradient:
 * logp [B,n] (optional), loss [B] = -logp[b, y_b] (optional), dlogits bf16 [B, dl_pitch] =
 * (softmax - onehot) * grad_scale (optional). */
int lbx_logsoftmax_xent(const float* logits, const int* labels, long long B, int n, float* logp, float* loss,
                        void* dlogits_bf16, int dl_pitch, float grad_scale, float* dbias /* optional: += column sums of
                        the gradient = bias gradient of the output layer */, void* stream);

/* Fused training head for few classes (N <= 8): Dense(K -> N) (lidbox/models/xvector.py:64) + log_softmax (:65) +
 * sparse cross-entropy (keras_utils.py:141-142) and the complete backward of that layer in one launch.
 * h_bf16 [B, K] (pitch ldh): output of the layer below; w_bf16 [K, ldw]: bf16 copy of the Keras kernel; bias [N] f32.
 * Outputs: loss [B]; logits_out [B, N] (optional); dh_bf16 [B, K] = d loss / d h (x grad_scale), zeroed where h <= 0
 * when relu_mask; dW [K, ldw], dbias [N], dbias_below [K] (optional) are ACCUMULATED (+=). */
int lbx_dense_xent_head(const void* h_bf16, const void* w_bf16, const float* bias, const int* labels, long long B, int K,
                        int N, int ldh, int ldw, float grad_scale, int relu_mask, float* logits_out, float* loss,
                        void* dh_bf16, float* dW, float* dbias, float* dbias_below, void* stream);

/* Fused dense head (lidbox/models/xvector.py:61-63: segment1 = Dense(K1 -> N1) + ReLU, segment2 = Dense(N1 -> N2) +
 * ReLU) in ONE persistent launch: h1 = bf16(relu(pooled . W1 + b1)), h2 = bf16(relu(h1 . W2 + b2)).  pooled [B, K1],
 * h1 [B, N1], h2 [B, N2] dense bf16; W1 [K1, ldw1], W2 [N1, ldw2] bf16 copies of the Keras kernels; scratch = scratch_floats
 * (>= B*N1, ideally 8*B*N1) floats of workspace — the split-K partial sums of segment1 go to one slab per split and are
 * added in a fixed order, so the result is bit-reproducible; sync_ws = 512 zero-initialised uint32 owned by the caller (grid-barrier
 * state: sync_ws[2] != 0 reports a barrier time-out).  All widths / pitches multiples of 8. */
int lbx_head_fwd(const void* pooled_bf16, long long B, int K1, const void* w1_bf16, int ldw1, const float* b1, int N1,
                 const void* w2_bf16, int ldw2, const float* b2, int N2, void* h1_bf16, void* h2_bf16, float* scratch,
                 long long scratch_floats, unsigned int* sync_ws, void* stream);
/* Backward of the same two layers in ONE launch, from dh2 = d loss / d (segment2 pre-activation) [B, N2] bf16:
 * dh1 = (dh2 . W2^T) * (h1 > 0) [B, N1] bf16, db1 += column sums of dh1, dW2 += h1^T . dh2, then
 * gpool = dh1 . W1^T [B, K1] fp32 (overwritten) and dW1 += pooled^T . dh1.  dW1 / dW2 have pitches ldw1 / ldw2. */
int lbx_head_bwd(const void* dh2_bf16, const void* pooled_bf16, const void* h1_bf16, long long B, int K1, int N1, int N2,
                 const void* w1_bf16, int ldw1, const void* w2_bf16, int ldw2, void* dh1_bf16, float* gpool, float* dw1,
                 float* db1, float* dw2, unsigned int* sync_ws, void* stream);

/* lidbox/losses.py:12-52 SparseAngularProximity(N, D, delta_weight): theta = acos(z[:, :N]),
 * loss[b] = sum_{l != y_b} sigmoid(delta_weight * (theta[b,y_b] - theta[b,l])).  normalize = 1 first maps
 * z = h / |h| (the L2-normalising head of the AP training config).  Optional outputs: z_out [B,D], theta_out [B,N],
 * loss [B], and the gradient w.r.t. h as f32 and/or bf16 [B, g_pitch], multiplied by gloss[b] (or 1) * grad_scale.
 * Constructor asserts of losses.py:14-16 are returned as LBX_EINVAL. */
int lbx_ap_loss(const float* h, const int* labels, long long B, int D, int N, float delta_weight, int normalize,
                float* z_out, float* theta_out, float* loss, float* grad_f32, void* grad_bf16, int g_pitch,
                const float* gloss, float grad_scale, float* dbias, void* stream);

/* Keras-compatible Adam on flat fp32 buffers: g' = g * grad_scale; m,v updates;
 * p -= lr * sqrt(1-beta2^t)/(1-beta1^t) * m / (sqrt(v) + eps).  The step counter t (*step_dev, incremented by the
 * call) and the bias-corrected rate (*lr_t_dev) live in device memory so the call can be replayed from a CUDA graph.
 * n must be a multiple of 4 and the buffers 16-byte aligned (vectorised). */
int lbx_adam_step(float* params, float* grads, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, long long* step_dev, float* lr_t_dev, float grad_scale,
                  void* params_bf16 /* optional: bf16 operand copy of the flat buffer, refreshed in the same pass */,
                  int zero_grads /* reset the gradient buffer after it has been consumed */, void* stream);

/* Data-parallel optimizer step fused with the gradient exchange over NVLink peer memory (one node, one process per
 * GPU; replaces "all-reduce + Adam").  The flat parameter / gradient / bf16-copy buffers are SYMMETRIC allocations:
 * params_ptrs / grads_ptrs / w16_ptrs / signal_ptrs are DEVICE arrays of `world` peer pointers (index = rank).
 * Every rank: waits until all ranks finished their backward pass (ONE peer flag exchange), sums ITS shard
 * [rank*n/world, (rank+1)*n/world) of all ranks' gradients with peer loads (reduce-scatter), applies Adam to the shard
 * (m_shard / v_shard hold n/world elements), stores the updated bf16 copy (and optionally the fp32 parameters) into
 * every rank's buffers with peer stores (all-gather) and publishes "pushed" to every peer — it does NOT wait for the
 * peers' pushes and does NOT touch the gradient buffer: the next step starts with lbx_dp_wait() (all peers have
 * published <=> this rank's bf16 weights are complete and nobody reads its gradient any more), after which the caller
 * clears its gradient buffer (off the critical path, e.g. on a side stream during the forward pass).
 * signal pads: 3*world uint32 per rank, zero-initialised; epoch_dev (1 uint32) / local_sync_dev (16 uint32; [4..9] receive three 64-bit globaltimer stamps of the last call:
 * start, barrier passed, last block done): zero-initialised local
 * state advanced by every call, so the calls can be replayed from a CUDA graph.  local_sync_dev[3] != 0 reports a
 * barrier time-out (lbx_set_dp_spin_limit polls of 64 ns, default 2^23 ~ 1 s): the update of that step is SKIPPED on
 * the rank that timed out, never applied from a half-reduced gradient; the host must treat it as fatal.
 * n must be a multiple of 4*world.  push_fp32 = 0 all-gathers only the bf16 operand copy (the fp32 master copy of a
 * shard then lives on its owner only, ZeRO-1 style; gather it from the peers before exporting).
 * mc_grads / mc_w16 (optional, both or none): NVLS multicast mappings of the gradient and bf16 buffers; when given the
 * reduce-scatter is one multimem.ld_reduce per element (the sum is formed inside the NVSwitch) and the all-gather one
 * multimem.st per element instead of `world` peer accesses. */
int lbx_adam_step_sharded(void* const* params_ptrs, void* const* grads_ptrs, void* const* w16_ptrs,
                          void* const* signal_ptrs, float* m_shard, float* v_shard, long long n, int rank, int world,
                          unsigned int* epoch_dev, unsigned int* local_sync_dev, float lr, float beta1, float beta2,
                          float eps, long long* step_dev, float* lr_t_dev, float grad_scale, int push_fp32,
                          const void* mc_grads, void* mc_w16, const float* staging, long long early_begin, void* stream);
/* Early part of the exchange, off the critical path: gradients at flat index >= early_begin (the later layers) are
 * complete long before the backward pass ends.  A rank announces that with lbx_dp_signal (slot_base = 2*world, epoch_add =
 * 1), a side stream waits for every peer's announcement with lbx_dp_wait_slot and then lets the COPY ENGINE pull the
 * peers' copies of this rank's shard (indices >= early_begin) into `staging` ((world-1) slabs of n/world floats, slab s
 * = peer (rank+1+s) % world) while the tensor cores finish the backward pass; lbx_adam_step_sharded then sums those
 * slabs locally and reads only the late part of the shard over NVLink.  staging = NULL: everything is read from the
 * peers inside the kernel. */
int lbx_dp_signal(void* const* signal_ptrs, int world, int rank, int slot_base, const unsigned int* epoch_dev,
                  unsigned int epoch_add, void* stream);
int lbx_dp_wait_slot(const void* signal_pad_local, int world, int slot_base, const unsigned int* epoch_dev,
                     unsigned int epoch_add, unsigned int* local_sync_dev, void* stream);
/* see lbx_adam_step_sharded: spins until every peer has published the epoch in *epoch_dev into this rank's pad */
int lbx_dp_wait(const void* signal_pad_local, int world, const unsigned int* epoch_dev, unsigned int* local_sync_dev,
                void* stream);
int lbx_set_dp_spin_limit(long long polls);
/* resident blocks per SM of lbx_adam_step_sharded (1..8, default 4) */
int lbx_set_dp_blocks_per_sm(int n);

/* fp32 -> bf16 operand planes of the flat parameter buffer: hi = bf16(x), lo = bf16(x - hi) (lo optional; it feeds
 * the bf16x3 forward mode).  The buffers keep the Keras layouts ([k*C_in, C_out] per kernel, pitch padded to 8). */
int lbx_split_bf16(const float* x, long long n, void* hi, void* lo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDBOX_B200_H */
