// Feature normalisation that directly follows log-mel in every lidbox pipeline:
// lidbox/features/__init__.py:5-9 feature_scaling, :16-20 cmn, :26-32 cmvn, :40-67 window_normalization.
// A tensor reduced over one axis is viewed as [outer, R, inner]; all statistics are fp32, two-pass (like TF).
#include "common.cuh"
#include <math.h>
#include <float.h>

namespace lbx {

__device__ __forceinline__ float div_no_nan(float a, float b) { return b == 0.0f ? 0.0f : a / b; }

// mode 0: cmn (x - mean), 1: cmvn ((x - mean) / std), 2: feature scaling lo + (hi-lo) * (x-min)/(max-min)
// one thread per (outer, inner) column; consecutive threads = consecutive inner -> coalesced when inner > 1
__global__ void __launch_bounds__(256) norm_columns_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          long long outer, long long R, long long inner, int mode,
                                                          float lo, float hi) {
  LBX_PDL_SYNC();
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= outer * inner) return;
  const long long o = col / inner, i = col - o * inner;
  const float* px = x + o * R * inner + i;
  float* py = y + o * R * inner + i;
  if (mode == 2) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (long long r = 0; r < R; ++r) {
      const float v = px[r * inner];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    const float range = mx - mn;
    for (long long r = 0; r < R; ++r) py[r * inner] = lo + (hi - lo) * div_no_nan(px[r * inner] - mn, range);
    return;
  }
  float s = 0.0f;
  for (long long r = 0; r < R; ++r) s += px[r * inner];
  const float mean = s / (float)R;
  float sd = 1.0f;
  if (mode == 1) {
    float q = 0.0f;
    for (long long r = 0; r < R; ++r) {
      const float d = px[r * inner] - mean;
      q = fmaf(d, d, q);
    }
    sd = sqrtf(q / (float)R);                       // tf.math.reduce_std: population
  }
  for (long long r = 0; r < R; ++r) {
    const float c = px[r * inner] - mean;
    py[r * inner] = mode == 1 ? div_no_nan(c, sd) : c;
  }
}

// inner == 1: one warp per row of length R
__global__ void __launch_bounds__(256) norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows,
                                                       long long R, int mode, float lo, float hi) {
  LBX_PDL_SYNC();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* px = x + row * R;
  float* py = y + row * R;
  if (mode == 2) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (long long r = lane; r < R; r += 32) {
      mn = fminf(mn, px[r]);
      mx = fmaxf(mx, px[r]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const float range = mx - mn;
    for (long long r = lane; r < R; r += 32) py[r] = lo + (hi - lo) * div_no_nan(px[r] - mn, range);
    return;
  }
  float s = 0.0f;
  for (long long r = lane; r < R; r += 32) s += px[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)R;
  float sd = 1.0f;
  if (mode == 1) {
    float q = 0.0f;
    for (long long r = lane; r < R; r += 32) {
      const float d = px[r] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    sd = sqrtf(q / (float)R);
  }
  for (long long r = lane; r < R; r += 32) {
    const float c = px[r] - mean;
    py[r] = mode == 1 ? div_no_nan(c, sd) : c;
  }
}

// global min/max (feature_scaling with axis=None): ordered-int atomics on a 2-float workspace
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void minmax_init_kernel(int* ws) {
  ws[0] = float_to_ordered(FLT_MAX);
  ws[1] = float_to_ordered(-FLT_MAX);
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, long long n, int* ws) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    mn = fminf(mn, x[i]);
    mx = fmaxf(mx, x[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(ws, float_to_ordered(mn));
    atomicMax(ws + 1, float_to_ordered(mx));
  }
}
__global__ void __launch_bounds__(256) scale_all_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                       const int* __restrict__ ws, float lo, float hi) {
  const float mn = ordered_to_float(ws[0]), range = ordered_to_float(ws[1]) - mn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = lo + (hi - lo) * div_no_nan(x[i] - mn, range);
}

// window_normalization over the time axis of [B, T, F] with REFLECT padding (features/__init__.py:40-67, T > window_len):
// left pad = w/2, right pad = w/2 - 1 + (w & 1); the window of output t covers padded indices [t, t + w)
__global__ void __launch_bounds__(256) window_norm_kernel(const float* __restrict__ x, float* __restrict__ y, long long B,
                                                         int T, int F, int w, int normalize_variance) {
  LBX_PDL_SYNC();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (b, t, f), f fastest
  if (idx >= B * T * F) return;
  const int f = (int)(idx % F);
  const long long bt = idx / F;
  const int t = (int)(bt % T);
  const long long b = bt / T;
  const float* px = x + b * T * (long long)F + f;
  const int left = w / 2;
  float s = 0.0f;
  for (int j = 0; j < w; ++j) {
    int o = t + j - left;
    o = o < 0 ? -o : (o >= T ? 2 * (T - 1) - o : o);
    s += px[(long long)o * F];
  }
  const float mean = s / (float)w;
  float out = px[(long long)t * F] - mean;
  if (normalize_variance) {
    float q = 0.0f;
    for (int j = 0; j < w; ++j) {
      int o = t + j - left;
      o = o < 0 ? -o : (o >= T ? 2 * (T - 1) - o : o);
      const float d = px[(long long)o * F] - mean;
      q = fmaf(d, d, q);
    }
    out = div_no_nan(out, sqrtf(q / (float)w));
  }
  y[idx] = out;
}

// tf.signal.mfccs_from_log_mel_spectrograms (call site lidbox/data/tf_utils.py:180-185):
// mfcc[k] = rsqrt(2 M) * 2 * sum_n x[n] cos(pi k (2n + 1) / (2 M)), k in [k0, k1)
__global__ void __launch_bounds__(256) mfcc_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows,
                                                  int M, int k0, int k1) {
  LBX_PDL_SYNC();
  const int nk = k1 - k0;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * nk) return;
  const long long r = idx / nk;
  const int k = k0 + (int)(idx - r * nk);
  const float* px = x + r * M;
  float acc = 0.0f;
  for (int n = 0; n < M; ++n) acc = fmaf(px[n], cospif((float)(k * (2 * n + 1)) / (float)(2 * M)), acc);
  y[idx] = 2.0f * acc * rsqrtf(2.0f * (float)M);
}

}  // namespace lbx

using namespace lbx;

extern "C" {

int lbx_normalize_axis_f32(const float* x, float* y, long long outer, long long R, long long inner, int mode, float lo,
                           float hi, void* stream) {
  LBX_CHECK_ARG(outer >= 0 && R >= 0 && inner >= 0, "negative extent");
  LBX_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0 (cmn), 1 (cmvn) or 2 (feature scaling)");
  if (outer * R * inner == 0) return LBX_OK;
  LBX_CHECK_ARG(x && y, "NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (inner == 1) {
    LBX_LAUNCH_PDL(norm_rows_kernel, dim3((unsigned)ceil_div(outer, 8)), dim3(256), 0, st, x, y, outer, R, mode, lo, hi);
  } else {
    LBX_LAUNCH_PDL(norm_columns_kernel, dim3((unsigned)ceil_div(outer * inner, 256)), dim3(256), 0, st, x, y, outer, R,
                   inner, mode, lo, hi);
  }
  return LBX_OK;
}

int lbx_feature_scaling_all_f32(const float* x, float* y, long long n, float lo, float hi, void* workspace,
                                void* stream) {
  LBX_CHECK_ARG(n >= 0, "negative length");
  if (n == 0) return LBX_OK;
  LBX_CHECK_ARG(x && y && workspace, "NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)(ceil_div(n, 256) < 148 * 8 ? ceil_div(n, 256) : 148 * 8);
  minmax_init_kernel<<<1, 1, 0, st>>>((int*)workspace);
  LBX_LAUNCH_CHECK();
  minmax_kernel<<<blocks, 256, 0, st>>>(x, n, (int*)workspace);
  LBX_LAUNCH_CHECK();
  scale_all_kernel<<<blocks, 256, 0, st>>>(x, y, n, (const int*)workspace, lo, hi);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

int lbx_mfcc_f32(const float* logmel, long long rows, int n_mel, int coef_begin, int coef_end, float* out, void* stream) {
  LBX_CHECK_ARG(rows >= 0 && n_mel >= 1, "bad shape");
  LBX_CHECK_ARG(coef_begin >= 0 && coef_end >= coef_begin && coef_end <= n_mel, "bad coefficient range [%d, %d)",
                coef_begin, coef_end);
  if (rows == 0 || coef_end == coef_begin) return LBX_OK;
  LBX_CHECK_ARG(logmel && out, "NULL pointer argument");
  const long long total = rows * (coef_end - coef_begin);
  LBX_LAUNCH_PDL(mfcc_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, (cudaStream_t)stream, logmel, out, rows,
                 n_mel, coef_begin, coef_end);
  return LBX_OK;
}

int lbx_window_normalization_f32(const float* x, float* y, long long B, int T, int F, int window_len,
                                 int normalize_variance, void* stream) {
  LBX_CHECK_ARG(B >= 0 && T >= 0 && F >= 0, "negative extent");
  LBX_CHECK_ARG(window_len >= 1 && window_len < T, "window_len must satisfy 1 <= window_len < T (otherwise use cmvn/cmn)");
  if (B * T * F == 0) return LBX_OK;
  LBX_CHECK_ARG(x && y, "NULL pointer argument");
  LBX_LAUNCH_PDL(window_norm_kernel, dim3((unsigned)ceil_div(B * T * (long long)F, 256)), dim3(256), 0,
                 (cudaStream_t)stream, x, y, B, T, F, window_len, normalize_variance);
  return LBX_OK;
}

}  // extern "C"
