// Non-GEMM pieces of the x-vector TDNN (lidbox/models/xvector.py) and its losses (lidbox/losses.py), sm_100a.
// All of them are HBM- or latency-bound: plain coalesced CUDA kernels, fp32 statistics.
#include "common.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

namespace lbx {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// counter-based hash RNG (one draw per (sample, channel)) for SpatialDropout1D
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned int a, unsigned int b) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * ((unsigned long long)a * 0x100000001B3ULL + b + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// advances a device-side call counter (SpatialDropout1D mask index): a kernel, so that CUDA-graph replays advance it too
__global__ void counter_tick_kernel(unsigned long long* counter) {
  LBX_PDL_SYNC();
  *counter += 1ULL;
}

// ------------------------------------------------------------------------------------------------------------
// features [B,T,F] f32 -> zero-left-padded bf16 activation rows (hi [+ lo]); optional SpatialDropout1D
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ x, long long B, int T, int F,
                                                       bf16* __restrict__ hi, bf16* __restrict__ lo, int rows_per_utt,
                                                       int row_off, int pitch, float drop_rate,
                                                       unsigned long long seed,
                                                       const unsigned long long* __restrict__ seed_counter) {
  LBX_PDL_SYNC();
  if (seed_counter != nullptr) seed += 7919ULL * *seed_counter;
  const long long total = B * T * (long long)pitch;
  const float keep_scale = drop_rate > 0.0f ? 1.0f / (1.0f - drop_rate) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % pitch);
    const long long bt = i / pitch;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    float v = 0.0f;
    if (c < F) {
      v = __ldg(x + bt * F + c);
      if (drop_rate > 0.0f) v = hash_uniform(seed, (unsigned)b, (unsigned)c) < drop_rate ? 0.0f : v * keep_scale;
    }
    const long long o = (b * rows_per_utt + row_off + t) * pitch + c;
    const bf16 h = __float2bfloat16_rn(v);
    hi[o] = h;
    if (lo != nullptr) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// vectorised variant (pitch % 8 == 0, 16-byte aligned planes): a thread converts 8 consecutive features of one frame
// (index arithmetic once per 8 elements, one 16-byte store per plane)
__global__ void __launch_bounds__(256) pack_rows_vec_kernel(const float* __restrict__ x, long long B, int T, int F,
                                                           bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                           int rows_per_utt, int row_off, int pitch, float drop_rate,
                                                           unsigned long long seed,
                                                           const unsigned long long* __restrict__ seed_counter,
                                                           int x_vec) {
  LBX_PDL_SYNC();
  if (seed_counter != nullptr) seed += 7919ULL * *seed_counter;
  const int gpr = pitch >> 3;                       // 8-feature groups per row
  const long long total = B * T * (long long)gpr;
  const float keep_scale = drop_rate > 0.0f ? 1.0f / (1.0f - drop_rate) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gi = (int)(i % gpr);
    const long long bt = i / gpr;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const int c0 = gi * 8;
    float v[8];
    const float* src = x + bt * F + c0;
    if (x_vec && c0 + 8 <= F) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = c0 + j < F ? __ldg(src + j) : 0.0f;
    }
    if (drop_rate > 0.0f) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < F) v[j] = hash_uniform(seed, (unsigned)b, (unsigned)(c0 + j)) < drop_rate ? 0.0f : v[j] * keep_scale;
    }
    const long long o = (b * rows_per_utt + row_off + t) * pitch + c0;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bf16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
      h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      const bf16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
      const bf16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
      l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo != nullptr) *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// GlobalMeanStddevPooling1D (xvector.py:25-35): two-pass population variance in fp32, clip 1e-10, sqrt
// ------------------------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float load_act(const TIn* p);
template <>
__device__ __forceinline__ float load_act<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_act<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void __launch_bounds__(256) stats_pool_fwd_kernel(const TIn* __restrict__ y, int rows_per_utt, int T, int C,
                                                            int pitch, float clip_min, float* __restrict__ out,
                                                            float* __restrict__ var_raw, bf16* __restrict__ out_hi,
                                                            bf16* __restrict__ out_lo) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long b = blockIdx.y;
  const TIn* base = y + b * rows_per_utt * (long long)pitch + c;
  float s = 0.0f;
  if (c < C)
    for (int t = tl; t < T; t += 8) s += load_act<TIn>(base + (long long)t * pitch);
  red[tl][cl] = s;
  __syncthreads();
  float mean = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mean += red[i][cl];
  mean /= (float)T;
  __syncthreads();
  float q = 0.0f;
  if (c < C)
    for (int t = tl; t < T; t += 8) {
      const float d = load_act<TIn>(base + (long long)t * pitch) - mean;
      q = fmaf(d, d, q);
    }
  red[tl][cl] = q;
  __syncthreads();
  if (tl == 0 && c < C) {
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += red[i][cl];
    var /= (float)T;
    const float sd = sqrtf(fminf(fmaxf(var, clip_min), 3.402823466e+38f));
    out[b * 2 * C + c] = mean;
    out[b * 2 * C + C + c] = sd;
    if (var_raw) var_raw[b * C + c] = var;
    if (out_hi) {
      const bf16 mh = __float2bfloat16_rn(mean), sh = __float2bfloat16_rn(sd);
      out_hi[b * 2 * C + c] = mh;
      out_hi[b * 2 * C + C + c] = sh;
      if (out_lo) {
        out_lo[b * 2 * C + c] = __float2bfloat16_rn(mean - __bfloat162float(mh));
        out_lo[b * 2 * C + C + c] = __float2bfloat16_rn(sd - __bfloat162float(sh));
      }
    }
  }
}

// backward of pooling fused with the ReLU mask of the producing frame layer:
//   dZ[b,t,c] = (y > 0) * ( g_mean/T + [var > clip] * g_std * (y - mean) / (T * std) )
__global__ void __launch_bounds__(256) stats_pool_bwd_kernel(const bf16* __restrict__ y, int rows_per_utt, int T, int C,
                                                            int pitch, float clip_min, const float* __restrict__ pooled,
                                                            const float* __restrict__ var_raw,
                                                            const float* __restrict__ gpool, bf16* __restrict__ dz) {
  const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long b = blockIdx.y;
  if (c >= C) return;
  const float mean = pooled[b * 2 * C + c], sd = pooled[b * 2 * C + C + c];
  const float gm = gpool[b * 2 * C + c] / (float)T;
  const float gs = var_raw[b * C + c] > clip_min ? gpool[b * 2 * C + C + c] / ((float)T * sd) : 0.0f;
  const long long base = b * rows_per_utt * (long long)pitch + c;
  for (int t = tl; t < T; t += 8) {
    const float v = __bfloat162float(y[base + (long long)t * pitch]);
    const float g = v > 0.0f ? fmaf(gs, v - mean, gm) : 0.0f;
    dz[base + (long long)t * pitch] = __float2bfloat16_rn(g);
  }
}

// ------------------------------------------------------------------------------------------------------------
// log_softmax (xvector.py:65) + sparse cross-entropy on the log-probs, forward and backward in one pass
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) logsoftmax_xent_kernel(const float* __restrict__ logits, const int* __restrict__ y,
                                                             long long B, int n, float* __restrict__ logp,
                                                             float* __restrict__ loss, bf16* __restrict__ dlogits,
                                                             int dl_pitch, float grad_scale, float* __restrict__ dbias) {
  LBX_PDL_SYNC();
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = logits + b * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float s = 0.0f;
  for (int j = lane; j < n; j += 32) s += expf(row[j] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  const int label = y ? y[b] : -1;
  // a label outside [0, n) raises in the reference (tf.gather / sparse cross-entropy); here the sample gets a NaN
  // loss and a zero gradient instead of a silent wrong value
  const bool bad_label = y != nullptr && (label < 0 || label >= n);
  if (bad_label && loss && lane == 0) loss[b] = __int_as_float(0x7fc00000);
  for (int j = lane; j < n; j += 32) {
    const float lp = row[j] - lse;
    if (logp) logp[b * n + j] = lp;
    if (dlogits) {
      const bf16 gq = __float2bfloat16_rn(bad_label ? 0.0f : (expf(lp) - (j == label ? 1.0f : 0.0f)) * grad_scale);
      dlogits[b * dl_pitch + j] = gq;
      if (dbias) atomicAdd(dbias + j, __bfloat162float(gq));
    }
    if (loss && j == label) loss[b] = -lp;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fused output layer of the training step for few classes (N <= 8): Dense(K -> N) (xvector.py:64) + log_softmax (:65) +
// sparse cross-entropy + the whole backward of that layer (data gradient masked by the ReLU of the layer below, weight
// and bias gradients, bias gradient of the layer below) in ONE launch instead of GEMM + loss + 2 GEMMs.  One warp per
// sample; weight-gradient partial sums of the 8 samples of a block are combined in shared memory before the atomics.
// ------------------------------------------------------------------------------------------------------------
constexpr int HEAD_NMAX = 8, HEAD_SPB = 8;

// KI = ceil(K / 32) rounded up to 16 or 32: a lane keeps its KI activations of the sample in registers (all loads in
// flight at once); the weights are staged once per block in shared memory as floats.
template <int KI>
__global__ void __launch_bounds__(32 * HEAD_SPB) dense_xent_head_kernel(
    const bf16* __restrict__ H, const bf16* __restrict__ W, const float* __restrict__ bias, const int* __restrict__ y,
    long long B, int K, int N, int ldh, int ldw, float grad_scale, int relu_mask, float* __restrict__ logits_out,
    float* __restrict__ loss, bf16* __restrict__ dH, float* __restrict__ dW, float* __restrict__ dbias,
    float* __restrict__ dbias_below) {
  LBX_PDL_SYNC();
  extern __shared__ float s_head[];
  float* s_w = s_head;                                   // [K][N] weights
  float* s_dbb = s_w + (size_t)K * N;                    // [K] bias gradient of the layer below (this block's samples)
  float* s_dl = s_dbb + K;                               // [SPB][NMAX] gradient w.r.t. the logits
  bf16* s_h = reinterpret_cast<bf16*>(s_dl + HEAD_SPB * HEAD_NMAX);   // [SPB][K] activations
#pragma unroll 8
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) s_w[i] = __bfloat162float(W[(long long)(i / N) * ldw + (i % N)]);
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_dbb[i] = 0.0f;
  if (threadIdx.x < HEAD_SPB * HEAD_NMAX) s_dl[threadIdx.x] = 0.0f;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * HEAD_SPB + wi;
  float hv[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const int h = i * 32 + lane;
    const bf16 raw = (b < B && h < K) ? H[b * ldh + h] : __float2bfloat16_rn(0.0f);
    hv[i] = __bfloat162float(raw);
    if (h < K) s_h[wi * K + h] = raw;
  }
  __syncthreads();
  if (b < B) {
    float acc[HEAD_NMAX];
#pragma unroll
    for (int j = 0; j < HEAD_NMAX; ++j) acc[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int h = i * 32 + lane;
      if (h < K) {
#pragma unroll
        for (int j = 0; j < HEAD_NMAX; ++j)
          if (j < N) acc[j] = fmaf(hv[i], s_w[h * N + j], acc[j]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < HEAD_NMAX; ++j) {
      if (j < N) {
        acc[j] = warp_sum(acc[j]) + bias[j];
        mx = fmaxf(mx, acc[j]);
      }
    }
    float se = 0.0f;
#pragma unroll
    for (int j = 0; j < HEAD_NMAX; ++j)
      if (j < N) se += expf(acc[j] - mx);
    const float lse = mx + logf(se);
    const int label = y[b];
    float dl[HEAD_NMAX];
#pragma unroll
    for (int j = 0; j < HEAD_NMAX; ++j) {
      dl[j] = 0.0f;
      if (j < N) {
        const float lp = acc[j] - lse;
        // the gradient w.r.t. the logits is stored / consumed in bf16, like every other data gradient of the path
        dl[j] = __bfloat162float(__float2bfloat16_rn((expf(lp) - (j == label ? 1.0f : 0.0f)) * grad_scale));
        if (lane == 0) {
          if (logits_out) logits_out[b * N + j] = acc[j];
          if (j == label) loss[b] = -lp;
          s_dl[wi * HEAD_NMAX + j] = dl[j];
        }
      }
    }
    bf16* drow = dH + b * ldh;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int h = i * 32 + lane;
      if (h < K) {
        float g = 0.0f;
#pragma unroll
        for (int j = 0; j < HEAD_NMAX; ++j)
          if (j < N) g = fmaf(dl[j], s_w[h * N + j], g);
        if (relu_mask && !(hv[i] > 0.0f)) g = 0.0f;
        const bf16 gq = __float2bfloat16_rn(g);
        drow[h] = gq;
        atomicAdd(s_dbb + h, __bfloat162float(gq));
      }
    }
  }
  __syncthreads();
  // weight gradient of this block's samples: every thread owns (h, j) pairs and sums over the samples in shared memory.
  // All blocks add into the same few cache lines, so the adds are issued as 16-byte vector reductions when the
  // layout allows (N == 4: one row of dW per instruction)
  if (N == 4 && (ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(dW) & 15) == 0) {
    for (int h = threadIdx.x; h < K; h += blockDim.x) {
      float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int sidx = 0; sidx < HEAD_SPB; ++sidx) {
        const float hs = __bfloat162float(s_h[sidx * K + h]);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaf(hs, s_dl[sidx * HEAD_NMAX + j], v[j]);
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dW + (long long)h * ldw), "f"(v[0]), "f"(v[1]),
                   "f"(v[2]), "f"(v[3])
                   : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
      const int h = i / N, j = i - h * N;
      float v = 0.0f;
#pragma unroll
      for (int sidx = 0; sidx < HEAD_SPB; ++sidx)
        v = fmaf(__bfloat162float(s_h[sidx * K + h]), s_dl[sidx * HEAD_NMAX + j], v);
      atomicAdd(dW + (long long)h * ldw + j, v);
    }
  }
  if (dbias_below != nullptr) {
    if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(dbias_below) & 15) == 0) {
      for (int h = 4 * threadIdx.x; h < K; h += 4 * blockDim.x)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbias_below + h), "f"(s_dbb[h]),
                     "f"(s_dbb[h + 1]), "f"(s_dbb[h + 2]), "f"(s_dbb[h + 3])
                     : "memory");
    } else {
      for (int h = threadIdx.x; h < K; h += blockDim.x) atomicAdd(dbias_below + h, s_dbb[h]);
    }
  }
  if (threadIdx.x < N) {
    float v = 0.0f;
#pragma unroll
    for (int sidx = 0; sidx < HEAD_SPB; ++sidx) v += s_dl[sidx * HEAD_NMAX + threadIdx.x];
    atomicAdd(dbias + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------------------------------------------------
// SparseAngularProximity (losses.py:12-52): theta = acos(z[:, :N]); L_b = sum_{l != y} sigmoid(w (theta_y - theta_l))
// optional L2-normalising head in front (z = h / |h|), as used by the AP training config
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ap_loss_kernel(const float* __restrict__ h, const int* __restrict__ y, long long B,
                                                     int D, int N, float w, int normalize, float* __restrict__ z_out,
                                                     float* __restrict__ theta_out, float* __restrict__ loss,
                                                     float* __restrict__ grad_f32, bf16* __restrict__ grad_bf16,
                                                     int g_pitch, const float* __restrict__ gloss, float grad_scale,
                                                     float* __restrict__ dbias) {
  LBX_PDL_SYNC();
  const long long b = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* row = h + b * D;
  float inv_norm = 1.0f;
  if (normalize) {
    float ss = 0.0f;
    for (int j = lane; j < D; j += 32) ss = fmaf(row[j], row[j], ss);
    ss = warp_sum(ss);
    inv_norm = rsqrtf(fmaxf(ss, 1e-12f));
  }
  int label = y[b];
  // labels outside [0, N) raise in the reference (tf.gather); here: NaN loss, zero gradient, no out-of-bounds read
  const bool bad_label = label < 0 || label >= N;
  if (bad_label) label = 0;
  const float zy = row[label] * inv_norm;
  const float theta_y = acosf(zy);
  float l = 0.0f, dty = 0.0f, dot = 0.0f;     // dot = z . dz (for the normalisation backward)
  for (int j0 = 0; j0 < D; j0 += 32) {
    const int j = j0 + lane;
    float dz = 0.0f, z = 0.0f;
    if (j < D) {
      z = row[j] * inv_norm;
      if (z_out) z_out[b * D + j] = z;
      if (j < N) {
        const float th = acosf(z);
        if (theta_out) theta_out[b * N + j] = th;
        if (j != label) {
          const float sg = 1.0f / (1.0f + expf(-w * (theta_y - th)));
          l += sg;
          const float dsg = w * sg * (1.0f - sg);          // d sigma / d delta * w
          dty += dsg;                                       // dL/dtheta_y
          dz = dsg * rsqrtf(fmaxf(1.0f - z * z, 0.0f));     // dL/dtheta_l = -dsg ; dtheta/dz = -1/sqrt(1-z^2)
        }
      }
    }
    dot += dz * z;
  }
  l = warp_sum(l);
  dty = warp_sum(dty);
  if (loss && lane == 0) loss[b] = bad_label ? __int_as_float(0x7fc00000) : l;
  if (grad_f32 == nullptr && grad_bf16 == nullptr) return;
  const float dzy = -dty * rsqrtf(fmaxf(1.0f - zy * zy, 0.0f));
  dot = warp_sum(dot) + dzy * zy;
  const float gl = bad_label ? 0.0f : (gloss ? gloss[b] : 1.0f) * grad_scale;
  for (int j = lane; j < D; j += 32) {
    const float z = row[j] * inv_norm;
    float dz = 0.0f;
    if (j < N) {
      if (j == label) {
        dz = dzy;
      } else {
        const float th = acosf(z);
        const float sg = 1.0f / (1.0f + expf(-w * (theta_y - th)));
        dz = w * sg * (1.0f - sg) * rsqrtf(fmaxf(1.0f - z * z, 0.0f));
      }
    }
    const float g = (normalize ? (dz - z * dot) * inv_norm : dz) * gl;
    if (grad_f32) grad_f32[b * g_pitch + j] = g;
    if (grad_bf16) {
      const bf16 gq = __float2bfloat16_rn(g);
      grad_bf16[b * g_pitch + j] = gq;
      if (dbias) atomicAdd(dbias + j, __bfloat162float(gq));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// optimizer + weight refresh
// ------------------------------------------------------------------------------------------------------------
// step counter and bias-corrected learning rate live in device memory so that a captured CUDA graph of the whole
// training step advances them on every replay
__global__ void adam_tick_kernel(long long* step, float* lr_t, float lr, float beta1, float beta2) {
  LBX_PDL_SYNC();
  const long long t = *step + 1;
  *step = t;
  *lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
}

// n4 = n / 4 float4 groups (the flat buffers are padded to a multiple of 4); optionally refreshes the bf16 operand
// copy of the parameters (same flat layout) and resets the gradient for the next step.  Every thread keeps
// 4 * LBX_ADAM_UNROLL 16-byte loads in flight; the moments are touched once per step and bypass L2 residency
// (ld/st .cs) so that they do not evict the activations and weights the next step re-reads.
#ifndef LBX_ADAM_UNROLL
#define LBX_ADAM_UNROLL 2
#endif
#ifndef LBX_ADAM_STREAM
#define LBX_ADAM_STREAM 1
#endif
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                  float4* __restrict__ v, long long n4,
                                                  const float* __restrict__ lr_t_ptr, float beta1, float beta2,
                                                  float eps, float grad_scale, uint2* __restrict__ p_bf16,
                                                  int zero_grads) {
  LBX_PDL_SYNC();
  constexpr int U = LBX_ADAM_UNROLL;
  const float lr_t = *lr_t_ptr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 gi[U], mi[U], vi[U], pi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        gi[u] = g[i];
        pi[u] = p[i];
#if LBX_ADAM_STREAM
        mi[u] = __ldcs(m + i);
        vi[u] = __ldcs(v + i);
#else
        mi[u] = m[i];
        vi[u] = v[i];
#endif
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) break;
      if (zero_grads) g[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#define LBX_ADAM1(c)                                              \
      {                                                           \
        const float gg = gi[u].c * grad_scale;                    \
        mi[u].c = beta1 * mi[u].c + (1.0f - beta1) * gg;          \
        vi[u].c = beta2 * vi[u].c + (1.0f - beta2) * gg * gg;     \
        pi[u].c -= lr_t * mi[u].c / (sqrtf(vi[u].c) + eps);       \
      }
      LBX_ADAM1(x) LBX_ADAM1(y) LBX_ADAM1(z) LBX_ADAM1(w)
#undef LBX_ADAM1
#if LBX_ADAM_STREAM
      __stcs(m + i, mi[u]);
      __stcs(v + i, vi[u]);
#else
      m[i] = mi[u];
      v[i] = vi[u];
#endif
      p[i] = pi[u];
      if (p_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pi[u].x, pi[u].y), hi = __floats2bfloat162_rn(pi[u].z, pi[u].w);
        p_bf16[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
}

// fp32 -> bf16 hi (+ lo residual) planes, elementwise (operand copies of the flat parameter buffer)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, long long n, bf16* __restrict__ hi,
                                                        bf16* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ------------------------------------------------------------------------------------------------------------
// vectorised bf16 pooling (training path): every thread owns 8 consecutive channels (one 16-byte load per row)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack_bf16x8(const uint4 u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256) stats_pool_fwd_bf16v_kernel(const bf16* __restrict__ y, int rows_per_utt, int T,
                                                                  int C, int pitch, float clip_min,
                                                                  float* __restrict__ out, float* __restrict__ var_raw,
                                                                  bf16* __restrict__ out_hi) {
  LBX_PDL_SYNC();
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int cv = blockIdx.x * 32 + lane;                    // 8-channel vector index
  const int c0 = cv * 8;
  const long long b = blockIdx.y;
  const bool active = c0 < pitch;
  const uint4* base = reinterpret_cast<const uint4*>(y + b * rows_per_utt * (long long)pitch) + cv;
  const int p8 = pitch >> 3;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
#pragma unroll 4
    for (int t = tl; t < T; t += 8) {
      float f[8];
      unpack_bf16x8(__ldg(base + (long long)t * p8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[tl][lane][i] = s[i];
  __syncthreads();
  float mean[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += red[j][lane][i];
    mean[i] = a / (float)T;
  }
  __syncthreads();
  float q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
#pragma unroll 4
    for (int t = tl; t < T; t += 8) {
      float f[8];
      unpack_bf16x8(__ldg(base + (long long)t * p8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = f[i] - mean[i];
        q[i] = fmaf(d, d, q[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[tl][lane][i] = q[i];
  __syncthreads();
  if (tl == 0 && active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c >= C) break;
      float var = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) var += red[j][lane][i];
      var /= (float)T;
      const float sd = sqrtf(fminf(fmaxf(var, clip_min), 3.402823466e+38f));
      out[b * 2 * C + c] = mean[i];
      out[b * 2 * C + C + c] = sd;
      if (var_raw) var_raw[b * C + c] = var;
      if (out_hi) {
        out_hi[b * 2 * C + c] = __float2bfloat16_rn(mean[i]);
        out_hi[b * 2 * C + C + c] = __float2bfloat16_rn(sd);
      }
    }
  }
}

// backward (+ ReLU mask) with the bias gradient of the producing layer accumulated on the fly; consumes (zeroes) gpool
__global__ void __launch_bounds__(256) stats_pool_bwd_bf16v_kernel(const bf16* __restrict__ y, int rows_per_utt, int T,
                                                                  int C, int pitch, float clip_min,
                                                                  const float* __restrict__ pooled,
                                                                  const float* __restrict__ var_raw,
                                                                  float* __restrict__ gpool, bf16* __restrict__ dz,
                                                                  float* __restrict__ dbias, int zero_gpool) {
  LBX_PDL_SYNC();
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int cv = blockIdx.x * 32 + lane;
  const int c0 = cv * 8;
  const long long b = blockIdx.y;
  const bool active = c0 < pitch;
  const int p8 = pitch >> 3;
  float mean[8], gm[8], gs[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    acc[i] = 0.0f;
    if (active && c < C) {
      mean[i] = pooled[b * 2 * C + c];
      const float sd = pooled[b * 2 * C + C + c];
      gm[i] = gpool[b * 2 * C + c] / (float)T;
      gs[i] = var_raw[b * C + c] > clip_min ? gpool[b * 2 * C + C + c] / ((float)T * sd) : 0.0f;
    } else {
      mean[i] = gm[i] = gs[i] = 0.0f;
    }
  }
  if (active) {
    const uint4* src = reinterpret_cast<const uint4*>(y + b * rows_per_utt * (long long)pitch) + cv;
    uint4* dst = reinterpret_cast<uint4*>(dz + b * rows_per_utt * (long long)pitch) + cv;
#pragma unroll 4
    for (int t = tl; t < T; t += 8) {
      float f[8], g[8];
      unpack_bf16x8(__ldg(src + (long long)t * p8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g[i] = f[i] > 0.0f ? fmaf(gs[i], f[i] - mean[i], gm[i]) : 0.0f;
        acc[i] += g[i];
      }
      dst[(long long)t * p8] = make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[tl][lane][i] = acc[i];
  }
  __syncthreads();
  if (tl == 0 && active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c >= C) break;
      if (dbias != nullptr) {
        float a = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += red[j][lane][i];
        atomicAdd(dbias + c, a);
      }
      if (zero_gpool) {
        gpool[b * 2 * C + c] = 0.0f;
        gpool[b * 2 * C + C + c] = 0.0f;
      }
    }
  }
}


// Column-owner variants for short time axes (T <= MAXT): a thread owns TWO channels of one utterance and keeps all T
// values in registers (one 4-byte load per row, a warp reads 128 contiguous bytes per row, every load is issued before
// the first use).  No cross-thread reduction is needed for the statistics; the backward kernel reduces the bias
// gradient over the 4 utterances of a block in shared memory before its atomics.
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

template <int MAXT>
__global__ void __launch_bounds__(128) stats_pool_fwd_bf16c_kernel(const bf16* __restrict__ y, int rows_per_utt, int T,
                                                                  int C, int pitch, float clip_min,
                                                                  float* __restrict__ out, float* __restrict__ var_raw,
                                                                  bf16* __restrict__ out_hi) {
  LBX_PDL_SYNC();
  const int cp = blockIdx.x * 128 + threadIdx.x;          // channel pair
  const int c0 = 2 * cp;
  if (c0 >= C) return;
  const long long b = blockIdx.y;
  const int p2 = pitch >> 1;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(y + b * rows_per_utt * (long long)pitch) + cp;
  uint32_t raw[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) raw[t] = t < T ? __ldg(base + (long long)t * p2) : 0u;
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {       // rows >= T hold zeros
    s0 += bf16lo(raw[t]);
    s1 += bf16hi(raw[t]);
  }
  const float inv_t = 1.0f / (float)T;
  const float m0 = s0 / (float)T, m1 = s1 / (float)T;
  float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t < T) {
      const float d0 = bf16lo(raw[t]) - m0, d1 = bf16hi(raw[t]) - m1;
      q0 = fmaf(d0, d0, q0);
      q1 = fmaf(d1, d1, q1);
    }
  }
  (void)inv_t;
  const float mean[2] = {m0, m1}, var[2] = {q0 / (float)T, q1 / (float)T};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = c0 + i;
    if (c >= C) break;
    const float sd = sqrtf(fminf(fmaxf(var[i], clip_min), 3.402823466e+38f));
    out[b * 2 * C + c] = mean[i];
    out[b * 2 * C + C + c] = sd;
    if (var_raw) var_raw[b * C + c] = var[i];
    if (out_hi) {
      out_hi[b * 2 * C + c] = __float2bfloat16_rn(mean[i]);
      out_hi[b * 2 * C + C + c] = __float2bfloat16_rn(sd);
    }
  }
}

template <int MAXT>
__global__ void __launch_bounds__(256) stats_pool_bwd_bf16c_kernel(const bf16* __restrict__ y, long long B,
                                                                  int rows_per_utt, int T, int C, int pitch,
                                                                  float clip_min, const float* __restrict__ pooled,
                                                                  const float* __restrict__ var_raw,
                                                                  float* __restrict__ gpool, bf16* __restrict__ dz,
                                                                  float* __restrict__ dbias, int zero_gpool) {
  LBX_PDL_SYNC();
  __shared__ float red[4][64][2];
  const int px = threadIdx.x & 63, bl = threadIdx.x >> 6;
  const int cp = blockIdx.x * 64 + px;
  const int c0 = 2 * cp;
  const long long b = (long long)blockIdx.y * 4 + bl;
  const bool active = c0 < C && b < B;
  float acc[2] = {0.0f, 0.0f};
  if (active) {
    const int p2 = pitch >> 1;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(y + b * rows_per_utt * (long long)pitch) + cp;
    uint32_t* dst = reinterpret_cast<uint32_t*>(dz + b * rows_per_utt * (long long)pitch) + cp;
    uint32_t raw[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) raw[t] = t < T ? __ldg(src + (long long)t * p2) : 0u;
    float mean[2], gm[2], gs[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = c0 + i;
      if (c < C) {
        mean[i] = pooled[b * 2 * C + c];
        const float sd = pooled[b * 2 * C + C + c];
        gm[i] = gpool[b * 2 * C + c] / (float)T;
        gs[i] = var_raw[b * C + c] > clip_min ? gpool[b * 2 * C + C + c] / ((float)T * sd) : 0.0f;
        if (zero_gpool) {
          gpool[b * 2 * C + c] = 0.0f;
          gpool[b * 2 * C + C + c] = 0.0f;
        }
      } else {
        mean[i] = gm[i] = gs[i] = 0.0f;
      }
    }
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      if (t < T) {
        const float f0 = bf16lo(raw[t]), f1 = bf16hi(raw[t]);
        const float g0 = f0 > 0.0f ? fmaf(gs[0], f0 - mean[0], gm[0]) : 0.0f;
        const float g1 = f1 > 0.0f ? fmaf(gs[1], f1 - mean[1], gm[1]) : 0.0f;
        acc[0] += g0;
        acc[1] += g1;
        dst[(long long)t * p2] = pack2(g0, g1);
      }
    }
  }
  if (dbias == nullptr) return;
  red[bl][px][0] = acc[0];
  red[bl][px][1] = acc[1];
  __syncthreads();
  if (bl == 0 && c0 < C) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (c0 + i < C) atomicAdd(dbias + c0 + i, red[0][px][i] + red[1][px][i] + red[2][px][i] + red[3][px][i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Sharded optimizer step fused with the gradient exchange over NVLink peer memory (data parallel, one node):
//   barrier -> reduce-scatter (peer loads of every rank's gradient shard) -> Adam on the shard -> all-gather (peer
//   stores of the updated fp32 parameters and bf16 operand copy into every rank) -> barrier -> local gradient reset.
// Replaces "NCCL all-reduce + full-size Adam": each rank moves 2/R of the gradient over NVLink instead of running a
// separate collective, and keeps Adam moments only for its own 1/R of the parameters.
// ------------------------------------------------------------------------------------------------------------
struct ShardedAdamParams {
  float* const* params;          // [world] peer pointers (symmetric buffers, identical layout)
  float* const* grads;
  bf16* const* w16;
  unsigned int* const* signals;  // [world] peer pointers to signal pads: 2*world uint32 each
  float* m;                      // local moments of the shard [shard_n]
  float* v;
  long long n;                   // flat length (multiple of 4*world)
  int rank, world;
  unsigned int* epoch;           // local: advanced by every call
  unsigned int* local_sync;      // local: [0] go1, [1] arrivals, [2] go2, [3] error flag
  long long* step;
  float* lr_t;
  float lr, beta1, beta2, eps, grad_scale;
  int push_fp32;                 // 1: also all-gather the fp32 master copy (otherwise only the owner's shard is current)
  const float* mc_grads;         // optional NVLS multicast mapping of the gradient buffers (in-switch reduction)
  bf16* mc_w16;                  // optional NVLS multicast mapping of the bf16 operand copy (one store reaches all ranks)
  const float* staging;          // optional: (world-1) slabs of n/world floats = the peers' copies of this rank's shard for
  long long early_begin;         //   flat indices >= early_begin, pulled by the copy engine during the backward pass
  long long spin_limit;          // polls (64 ns apart) before a barrier gives up and the step is skipped
  int debug;                     // measurement only: 1 = no remote loads, 2 = no remote stores, 4 = no fence per thread
};

// NVLink SHARP: one load returns the sum over all ranks of the multicast group, reduced inside the NVSwitch
__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_ptr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_b32x2(void* mc_ptr, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc_ptr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y))
               : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a lost peer must not hang the GPU (error flag instead)
__device__ __forceinline__ bool spin_until_ge(const unsigned int* p, unsigned int target, bool sys_scope,
                                              long long limit = (1LL << 23)) {
  for (long long it = 0; it < limit; ++it) {         // 2^23 polls ~ 1 s
    const unsigned int v = sys_scope ? ld_acquire_sys(p) : ld_acquire_gpu(p);
    if ((int)(v - target) >= 0) return true;
    __nanosleep(64);
  }
  return false;
}

// cross-GPU barrier executed by thread 0 of block 0: slot `base + rank` of every peer's pad receives the epoch
__device__ __forceinline__ bool peer_barrier(const ShardedAdamParams& p, unsigned int e, int base) {
  for (int q = 0; q < p.world; ++q) st_release_sys(p.signals[q] + base + p.rank, e);
  bool ok = true;
  for (int q = 0; q < p.world; ++q) ok = spin_until_ge(p.signals[p.rank] + base + q, e, true, p.spin_limit) && ok;
  return ok;
}

// NVLS: gradients are reduced inside the NVSwitch (one multimem.ld_reduce per element), W is unused (1);
// otherwise W = number of ranks whose gradient shard is read over NVLink peer mappings (2, 4 or 8).
//
// Critical path of a data-parallel step after the backward pass = ONE flag exchange (all ranks have finished their
// gradients) + the shard update itself.  Nothing else waits inside this kernel: every block leaves as soon as its part
// of the shard is pushed, the LAST block to finish publishes "rank r has pushed epoch e" to every peer, and the next
// step's forward pass starts with dp_wait_kernel (weights complete on this rank <=> all peers have published), after
// which the local gradient is cleared off the critical path.  A barrier time-out (a peer is more than spin_limit polls
// late) sets local_sync[3] and SKIPS the update on this rank: no half-reduced gradient is ever applied.
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool NVLS, int W>
__global__ void __launch_bounds__(256) adam_sharded_kernel(const ShardedAdamParams p) {
  LBX_PDL_SYNC();
  __shared__ unsigned int s_abort;
  __shared__ float s_lr_t;
  const unsigned int nblk = gridDim.x;
  const unsigned int e = *p.epoch + 1;            // epoch is only advanced at the very end, by the last block
  // optional phase timestamps (local_sync[4..9] as three 64-bit nanosecond stamps: start, barrier passed, last block done)
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(p.local_sync + 4);
  // ---- barrier: every rank has finished its backward pass (its gradient buffer is complete).  Block 0 announces this
  // rank to every peer; EVERY block polls this rank's own pad (local memory the peers write into): no relay hop ----
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      stamps[0] = global_timer_ns();
      __threadfence_system();                     // this rank's gradient writes are visible to the peers
      for (int q = 0; q < p.world; ++q) st_release_sys(p.signals[q] + p.rank, e);
    }
    bool ok = true;
    for (int q = 0; q < p.world; ++q) ok = spin_until_ge(p.signals[p.rank] + q, e, true, p.spin_limit) && ok;
    if (!ok) p.local_sync[3] = 1;
    if (blockIdx.x == 0) {
      stamps[1] = global_timer_ns();
      if (ok) {
        const long long t = *p.step + 1;
        *p.step = t;
      }
    }
    // bias-corrected rate of THIS step, computed by every block from the (not yet advanced or just advanced) counter
    const long long t = e;                        // one optimizer step per epoch: step counter == epoch
    s_lr_t = (float)((double)p.lr * sqrt(1.0 - pow((double)p.beta2, (double)t)) / (1.0 - pow((double)p.beta1, (double)t)));
    if (blockIdx.x == 0) *p.lr_t = s_lr_t;
    s_abort = ok ? 0u : 1u;
  }
  __syncthreads();
  const float lr_t = s_lr_t;

  // ---- reduce-scatter + Adam + all-gather on this rank's shard ----
  const long long shard4 = s_abort ? 0 : p.n / 4 / p.world;    // float4 groups per shard (none after a time-out)
  const long long base4 = (p.n / 4 / p.world) * p.rank;
  float4* m4 = reinterpret_cast<float4*>(p.m);
  float4* v4 = reinterpret_cast<float4*>(p.v);
  float4* my_params = reinterpret_cast<float4*>(p.params[p.rank]) + base4;
  // Software pipeline: the NVLink loads (gradient shards of the peers, or one in-switch reduction) of iteration k+1 are
  // in flight while iteration k does its local work (moments, master weights, bf16 push), so the link time and the HBM
  // time overlap instead of adding up (measured on 2 GPUs: 17 us of exposed peer-load latency with a single wave).
  constexpr int UNROLL = 2;
  const long long stride = (long long)nblk * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float4 tn[UNROLL][W];
  // elements of the shard at flat index >= early_begin: their gradients were complete long before the end of the
  // backward pass and the copy engine has already pulled the peers' copies into the local staging slabs, so they need
  // no NVLink round trip here — only the late part of the shard (the first frame layers) is read from the peers
  const long long shard4_full = p.n / 4 / p.world;
  const long long early4 = p.staging != nullptr ? p.early_begin / 4 : (1LL << 60);
  const float4* stg4 = reinterpret_cast<const float4*>(p.staging);
  const float4* my_grads4 = reinterpret_cast<const float4*>(p.grads[p.rank]);
  auto fetch = [&](long long i0) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < shard4) {
        if (base4 + i >= early4) {
          float4 acc = __ldcv(my_grads4 + base4 + i);
          for (int sl = 0; sl < p.world - 1; ++sl) {
            const float4 v = __ldcs(stg4 + sl * shard4_full + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          tn[u][0] = acc;
#pragma unroll
          for (int q = 1; q < W; ++q) tn[u][q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else if (NVLS) {
          tn[u][0] = multimem_ld_reduce_f32x4(p.mc_grads + 4 * (base4 + i));
        } else {
#pragma unroll
          for (int q = 0; q < W; ++q)
            tn[u][q] = __ldcv(reinterpret_cast<const float4*>(p.grads[(p.debug & 1) ? p.rank : q]) + base4 + i);
        }
      }
    }
  };
  fetch(first);
  for (long long i0 = first; i0 < shard4; i0 += stride * UNROLL) {
    float4 t[UNROLL][W], mi[UNROLL], vi[UNROLL], pi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int q = 0; q < W; ++q) t[u][q] = tn[u][q];
      const long long i = i0 + u * stride;
      if (i < shard4) {
        mi[u] = __ldcs(m4 + i);
        vi[u] = __ldcs(v4 + i);
        pi[u] = my_params[i];
      }
    }
    fetch(i0 + stride * UNROLL);                  // next iteration's link traffic, before this iteration's math / stores
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i >= shard4) break;
      float4 g = t[u][0];
      if (!NVLS) {
#pragma unroll
        for (int q = 1; q < W; ++q) { g.x += t[u][q].x; g.y += t[u][q].y; g.z += t[u][q].z; g.w += t[u][q].w; }
      }
#define LBX_ADAM1(c)                                                  \
      {                                                               \
        const float gg = g.c * p.grad_scale;                          \
        mi[u].c = p.beta1 * mi[u].c + (1.0f - p.beta1) * gg;          \
        vi[u].c = p.beta2 * vi[u].c + (1.0f - p.beta2) * gg * gg;     \
        pi[u].c -= lr_t * mi[u].c / (sqrtf(vi[u].c) + p.eps);         \
      }
      LBX_ADAM1(x) LBX_ADAM1(y) LBX_ADAM1(z) LBX_ADAM1(w)
#undef LBX_ADAM1
      __stcs(m4 + i, mi[u]);
      __stcs(v4 + i, vi[u]);
      __nv_bfloat162 lo = __floats2bfloat162_rn(pi[u].x, pi[u].y), hi = __floats2bfloat162_rn(pi[u].z, pi[u].w);
      const uint2 packed = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      if (NVLS && !p.push_fp32) {                   // all-gather: one multicast store reaches every rank
        my_params[i] = pi[u];
        multimem_st_b32x2(p.mc_w16 + 4 * (base4 + i), packed);
      } else {
        for (int q = 0; q < p.world; ++q) {         // all-gather by peer stores
          if ((p.debug & 2) && q != p.rank) continue;
          if (p.push_fp32 || q == p.rank) reinterpret_cast<float4*>(p.params[q])[base4 + i] = pi[u];
          reinterpret_cast<uint2*>(p.w16[q])[base4 + i] = packed;
        }
      }
    }
  }
  // ---- publish: the last block to finish tells every peer that this rank's shard has been pushed everywhere (and that
  // this rank no longer reads anybody's gradient); nobody waits here.  bar.sync + one system fence by thread 0 orders
  // the whole block's remote stores before the count (the cooperative-groups grid-sync pattern) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int done = atomicAdd(p.local_sync + 1, 1u) + 1u;
    if (done == e * nblk) {                       // counts accumulate over the epochs: e * nblk after epoch e
      stamps[2] = global_timer_ns();
      __threadfence_system();
      for (int q = 0; q < p.world; ++q) st_release_sys(p.signals[q] + p.world + p.rank, e);
      *p.epoch = e;                               // every block has read the old value (they all passed the start)
    }
  }
}

// Start of the next step: all peers have published epoch *epoch  <=>  every shard of the bf16 weights has landed in
// this rank's copy and nobody reads this rank's gradient buffer any more (it may be cleared).
__global__ void dp_wait_kernel(const unsigned int* __restrict__ pad, int world, const unsigned int* __restrict__ epoch,
                               unsigned int* __restrict__ local_sync, long long spin_limit) {
  LBX_PDL_SYNC();
  const unsigned int e = *epoch;
  if ((int)threadIdx.x < world && !spin_until_ge(pad + world + threadIdx.x, e, true, spin_limit)) local_sync[3] = 1;
}

// "rank r has reached point X of epoch *epoch + add": one flag per peer pad, slot_base + rank
__global__ void dp_signal_kernel(unsigned int* const* __restrict__ signals, int world, int rank, int slot_base,
                                 const unsigned int* __restrict__ epoch, unsigned int add) {
  LBX_PDL_SYNC();
  const unsigned int e = *epoch + add;
  if ((int)threadIdx.x < world) {
    __threadfence_system();                      // the gradients written by the kernels before this one are visible
    st_release_sys(signals[threadIdx.x] + slot_base + rank, e);
  }
}
__global__ void dp_wait_slot_kernel(const unsigned int* __restrict__ pad, int world, int slot_base,
                                    const unsigned int* __restrict__ epoch, unsigned int add,
                                    unsigned int* __restrict__ local_sync, long long spin_limit) {
  LBX_PDL_SYNC();
  const unsigned int e = *epoch + add;
  if ((int)threadIdx.x < world && !spin_until_ge(pad + slot_base + threadIdx.x, e, true, spin_limit)) local_sync[3] = 1;
}

static long long g_dp_spin_limit = 1LL << 23;      // ~1 s of 64 ns polls
static int g_dp_blocks_per_sm = 2;

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = ceil_div(n, block);
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace lbx

using namespace lbx;

extern "C" {

int lbx_counter_tick(unsigned long long* counter_dev, void* stream) {
  LBX_CHECK_ARG(counter_dev != nullptr, "NULL counter");
  LBX_LAUNCH_PDL(counter_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, counter_dev);
  return LBX_OK;
}

int lbx_pack_rows_bf16(const float* x, long long B, int T, int F, void* hi, void* lo, int rows_per_utt, int row_off,
                       int pitch, float drop_rate, unsigned long long seed, const unsigned long long* seed_counter_dev,
                       void* stream) {
  LBX_CHECK_ARG(B >= 0 && T >= 0 && F >= 1 && pitch >= F, "bad shape B=%lld T=%d F=%d pitch=%d", B, T, F, pitch);
  LBX_CHECK_ARG(row_off >= 0 && row_off + T <= rows_per_utt, "rows do not fit: off=%d T=%d rows_per_utt=%d", row_off, T,
                rows_per_utt);
  LBX_CHECK_ARG(drop_rate >= 0.0f && drop_rate < 1.0f, "drop_rate must be in [0, 1)");
  if (B * T == 0) return LBX_OK;
  LBX_CHECK_ARG(x && hi, "NULL pointer argument");
  if (pitch % 8 == 0 && ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
    const int x_vec = F % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    LBX_LAUNCH_PDL(pack_rows_vec_kernel, dim3(grid_for(B * T * (long long)(pitch / 8), 256)), dim3(256), 0,
                   (cudaStream_t)stream, x, B, T, F, (bf16*)hi, (bf16*)lo, rows_per_utt, row_off, pitch, drop_rate, seed,
                   seed_counter_dev, x_vec);
    return LBX_OK;
  }
  LBX_LAUNCH_PDL(pack_rows_kernel, dim3(grid_for(B * T * (long long)pitch, 256)), dim3(256), 0, (cudaStream_t)stream, x, B,
                 T, F, (bf16*)hi, (bf16*)lo, rows_per_utt, row_off, pitch, drop_rate, seed, seed_counter_dev);
  return LBX_OK;
}

int lbx_stats_pool_fwd(const void* y, int y_dtype, long long B, int rows_per_utt, int T, int C, int pitch,
                       float clip_min, float* out, float* var_raw, void* out_hi, void* out_lo, void* stream) {
  LBX_CHECK_ARG(B >= 0 && T >= 1 && C >= 1 && pitch >= C && rows_per_utt >= T, "bad pooling shape");
  LBX_CHECK_ARG(B <= 65535, "batch too large for one pooling launch");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(y && out, "NULL pointer argument");
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)B);
  if (y_dtype == LBX_F32)
    stats_pool_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)y, rows_per_utt, T, C, pitch,
                                                                         clip_min, out, var_raw, (bf16*)out_hi,
                                                                         (bf16*)out_lo);
  else if (y_dtype == LBX_BF16 && out_lo == nullptr && pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0)
  {
    if (T <= 56) {
      const dim3 g2((unsigned)ceil_div((C + 1) / 2, 128), (unsigned)B);
      if (T <= 40)
        LBX_LAUNCH_PDL(stats_pool_fwd_bf16c_kernel<40>, g2, dim3(128), 0, (cudaStream_t)stream, (const bf16*)y,
                       rows_per_utt, T, C, pitch, clip_min, out, var_raw, (bf16*)out_hi);
      else
        LBX_LAUNCH_PDL(stats_pool_fwd_bf16c_kernel<56>, g2, dim3(128), 0, (cudaStream_t)stream, (const bf16*)y,
                       rows_per_utt, T, C, pitch, clip_min, out, var_raw, (bf16*)out_hi);
      return LBX_OK;
    }
    LBX_LAUNCH_PDL(stats_pool_fwd_bf16v_kernel, dim3((unsigned)ceil_div(pitch / 8, 32), (unsigned)B), dim3(256), 0,
                   (cudaStream_t)stream, (const bf16*)y, rows_per_utt, T, C, pitch, clip_min, out, var_raw,
                   (bf16*)out_hi);
    return LBX_OK;
  }
  else if (y_dtype == LBX_BF16)
    stats_pool_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, rows_per_utt, T, C, pitch,
                                                                        clip_min, out, var_raw, (bf16*)out_hi,
                                                                        (bf16*)out_lo);
  else
    return set_error(LBX_EINVAL, "bad y_dtype %d", y_dtype);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

int lbx_stats_pool_bwd(const void* y_bf16, long long B, int rows_per_utt, int T, int C, int pitch, float clip_min,
                       const float* pooled, const float* var_raw, float* gpool, void* dz_bf16, float* dbias,
                       int zero_gpool, void* stream) {
  LBX_CHECK_ARG(B >= 0 && T >= 1 && C >= 1 && pitch >= C && rows_per_utt >= T && B <= 65535, "bad pooling shape");
  LBX_CHECK_ARG(pitch % 8 == 0, "pitch must be a multiple of 8");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(y_bf16 && pooled && var_raw && gpool && dz_bf16, "NULL pointer argument");
  LBX_CHECK_ARG(((reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(dz_bf16)) & 15) == 0,
                "activation buffers must be 16-byte aligned");
  dim3 grid((unsigned)ceil_div(pitch / 8, 32), (unsigned)B);
  if (T <= 56) {
    const dim3 g2((unsigned)ceil_div((C + 1) / 2, 64), (unsigned)ceil_div(B, 4));
    if (T <= 40)
      LBX_LAUNCH_PDL(stats_pool_bwd_bf16c_kernel<40>, g2, dim3(256), 0, (cudaStream_t)stream, (const bf16*)y_bf16, B,
                     rows_per_utt, T, C, pitch, clip_min, pooled, var_raw, gpool, (bf16*)dz_bf16, dbias, zero_gpool);
    else
      LBX_LAUNCH_PDL(stats_pool_bwd_bf16c_kernel<56>, g2, dim3(256), 0, (cudaStream_t)stream, (const bf16*)y_bf16, B,
                     rows_per_utt, T, C, pitch, clip_min, pooled, var_raw, gpool, (bf16*)dz_bf16, dbias, zero_gpool);
    return LBX_OK;
  }
  LBX_LAUNCH_PDL(stats_pool_bwd_bf16v_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const bf16*)y_bf16, rows_per_utt, T,
                 C, pitch, clip_min, pooled, var_raw, gpool, (bf16*)dz_bf16, dbias, zero_gpool);
  return LBX_OK;
}

int lbx_logsoftmax_xent(const float* logits, const int* labels, long long B, int n, float* logp, float* loss,
                        void* dlogits_bf16, int dl_pitch, float grad_scale, float* dbias, void* stream) {
  LBX_CHECK_ARG(B >= 0 && n >= 1, "bad shape");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(logits, "NULL logits");
  LBX_CHECK_ARG(!(loss || dlogits_bf16) || labels, "labels are required for the loss / gradient");
  LBX_CHECK_ARG(!dlogits_bf16 || dl_pitch >= n, "dl_pitch too small");
  LBX_LAUNCH_PDL(logsoftmax_xent_kernel, dim3((unsigned)ceil_div(B, 4)), dim3(128), 0, (cudaStream_t)stream, logits, labels,
                 B, n, logp, loss, (bf16*)dlogits_bf16, dl_pitch, grad_scale, dbias);
  return LBX_OK;
}

int lbx_dense_xent_head(const void* h_bf16, const void* w_bf16, const float* bias, const int* labels, long long B, int K,
                        int N, int ldh, int ldw, float grad_scale, int relu_mask, float* logits_out, float* loss,
                        void* dh_bf16, float* dW, float* dbias, float* dbias_below, void* stream) {
  LBX_CHECK_ARG(B >= 0 && K >= 1 && N >= 1 && N <= HEAD_NMAX && ldh >= K && ldw >= N, "bad shape (N <= %d)", HEAD_NMAX);
  LBX_CHECK_ARG(K <= 1024, "K must be <= 1024 for the fused head");
  const size_t smem = (size_t)(K * N + K + HEAD_SPB * HEAD_NMAX) * sizeof(float) + (size_t)HEAD_SPB * K * sizeof(bf16);
  LBX_CHECK_ARG(smem <= 48 * 1024, "K * N too large for the fused head");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(h_bf16 && w_bf16 && bias && labels && loss && dh_bf16 && dW && dbias, "NULL pointer argument");
  const dim3 grid((unsigned)ceil_div(B, HEAD_SPB)), block(32 * HEAD_SPB);
  if (K <= 512)
    LBX_LAUNCH_PDL(dense_xent_head_kernel<16>, grid, block, smem, (cudaStream_t)stream, (const bf16*)h_bf16,
                   (const bf16*)w_bf16, bias, labels, B, K, N, ldh, ldw, grad_scale, relu_mask, logits_out, loss,
                   (bf16*)dh_bf16, dW, dbias, dbias_below);
  else
    LBX_LAUNCH_PDL(dense_xent_head_kernel<32>, grid, block, smem, (cudaStream_t)stream, (const bf16*)h_bf16,
                   (const bf16*)w_bf16, bias, labels, B, K, N, ldh, ldw, grad_scale, relu_mask, logits_out, loss,
                   (bf16*)dh_bf16, dW, dbias, dbias_below);
  return LBX_OK;
}

int lbx_ap_loss(const float* h, const int* labels, long long B, int D, int N, float delta_weight, int normalize,
                float* z_out, float* theta_out, float* loss, float* grad_f32, void* grad_bf16, int g_pitch,
                const float* gloss, float grad_scale, float* dbias, void* stream) {
  LBX_CHECK_ARG(N >= 1, "Must have at least 1 class");                                       /* losses.py:14 */
  LBX_CHECK_ARG(D >= N, "Language vector dimension cannot be less than number of classes");  /* losses.py:15 */
  LBX_CHECK_ARG(delta_weight > 0.0f, "delta_weight must be positive");                       /* losses.py:16 */
  LBX_CHECK_ARG(B >= 0, "bad batch");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(h && labels, "NULL pointer argument");
  LBX_CHECK_ARG(!(grad_f32 || grad_bf16) || g_pitch >= D, "g_pitch too small");
  LBX_LAUNCH_PDL(ap_loss_kernel, dim3((unsigned)ceil_div(B, 4)), dim3(128), 0, (cudaStream_t)stream, h, labels, B, D, N,
                 delta_weight, normalize, z_out, theta_out, loss, grad_f32, (bf16*)grad_bf16, g_pitch, gloss,
                 grad_scale, dbias);
  return LBX_OK;
}

int lbx_adam_step(float* params, float* grads, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, long long* step_dev, float* lr_t_dev, float grad_scale, void* params_bf16, int zero_grads,
                  void* stream) {
  LBX_CHECK_ARG(n >= 0 && n % 4 == 0, "the flat parameter count must be a multiple of 4 (pad the buffers)");
  if (n == 0) return LBX_OK;
  LBX_CHECK_ARG(params && grads && m && v && step_dev && lr_t_dev, "NULL pointer argument");
  LBX_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                  reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(params_bf16) & 7) == 0,
                "flat buffers must be 16-byte aligned");
  LBX_LAUNCH_PDL(adam_tick_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev, lr_t_dev, lr, beta1, beta2);
  LBX_LAUNCH_PDL(adam_kernel, dim3(grid_for(ceil_div(n / 4, LBX_ADAM_UNROLL), 256)), dim3(256), 0, (cudaStream_t)stream, (float4*)params,
                 (float4*)grads, (float4*)m, (float4*)v, n / 4, (const float*)lr_t_dev, beta1, beta2, eps, grad_scale,
                 (uint2*)params_bf16, zero_grads);
  return LBX_OK;
}

int lbx_split_bf16(const float* x, long long n, void* hi, void* lo, void* stream) {
  LBX_CHECK_ARG(n >= 0, "bad length");
  if (n == 0) return LBX_OK;
  LBX_CHECK_ARG(x && hi, "NULL pointer argument");
  split_bf16_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, (bf16*)hi, (bf16*)lo);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

int lbx_adam_step_sharded(void* const* params_ptrs, void* const* grads_ptrs, void* const* w16_ptrs,
                          void* const* signal_ptrs, float* m_shard, float* v_shard, long long n, int rank, int world,
                          unsigned int* epoch_dev, unsigned int* local_sync_dev, float lr, float beta1, float beta2,
                          float eps, long long* step_dev, float* lr_t_dev, float grad_scale, int push_fp32,
                          const void* mc_grads, void* mc_w16, const float* staging, long long early_begin, void* stream) {
  LBX_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  LBX_CHECK_ARG(n > 0 && n % (4LL * world) == 0, "the flat length must be a multiple of 4*world (pad the buffers)");
  LBX_CHECK_ARG(params_ptrs && grads_ptrs && w16_ptrs && signal_ptrs && m_shard && v_shard && epoch_dev &&
                    local_sync_dev && step_dev && lr_t_dev,
                "NULL pointer argument");
  ShardedAdamParams p{};
  p.params = (float* const*)params_ptrs; p.grads = (float* const*)grads_ptrs; p.w16 = (bf16* const*)w16_ptrs;
  p.signals = (unsigned int* const*)signal_ptrs;
  p.m = m_shard; p.v = v_shard; p.n = n; p.rank = rank; p.world = world;
  p.epoch = epoch_dev; p.local_sync = local_sync_dev; p.step = step_dev; p.lr_t = lr_t_dev;
  p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.grad_scale = grad_scale;
  p.push_fp32 = push_fp32;
  p.mc_grads = (const float*)mc_grads; p.mc_w16 = (bf16*)mc_w16;
  LBX_CHECK_ARG(staging == nullptr || (early_begin >= 0 && early_begin % 4 == 0 &&
                                       (reinterpret_cast<uintptr_t>(staging) & 15) == 0),
                "staging needs a 16-byte aligned buffer and early_begin % 4 == 0");
  p.staging = staging; p.early_begin = early_begin;
  p.spin_limit = g_dp_spin_limit;
  p.debug = getenv("LBX_DP_DEBUG") ? atoi(getenv("LBX_DP_DEBUG")) : 0;
  LBX_CHECK_ARG(world <= 8, "at most 8 ranks (one NVLink domain)");
  // every block must be able to be resident at once (grid-wide flags): occupancy-limited grid
  int dev = 0, sms = 0, per_sm = 0;
  LBX_CUDA(cudaGetDevice(&dev));
  LBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void (*kernel)(const ShardedAdamParams) = nullptr;
  if (p.mc_grads != nullptr && p.mc_w16 != nullptr) {
    kernel = adam_sharded_kernel<true, 1>;
  } else {
    switch (world) {
      case 1: kernel = adam_sharded_kernel<false, 1>; break;
      case 2: kernel = adam_sharded_kernel<false, 2>; break;
      case 3: kernel = adam_sharded_kernel<false, 3>; break;
      case 4: kernel = adam_sharded_kernel<false, 4>; break;
      case 5: kernel = adam_sharded_kernel<false, 5>; break;
      case 6: kernel = adam_sharded_kernel<false, 6>; break;
      case 7: kernel = adam_sharded_kernel<false, 7>; break;
      default: kernel = adam_sharded_kernel<false, 8>; break;
    }
  }
  LBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0));
  if (per_sm > g_dp_blocks_per_sm) per_sm = g_dp_blocks_per_sm;
  if (per_sm < 1) per_sm = 1;
  LBX_LAUNCH_PDL(kernel, dim3((unsigned)(sms * per_sm)), dim3(256), 0, (cudaStream_t)stream, p);
  return LBX_OK;
}

int lbx_dp_wait(const void* signal_pad_local, int world, const unsigned int* epoch_dev, unsigned int* local_sync_dev,
                void* stream) {
  LBX_CHECK_ARG(signal_pad_local && epoch_dev && local_sync_dev, "NULL pointer argument");
  LBX_CHECK_ARG(world >= 1 && world <= 8, "at most 8 ranks (one NVLink domain)");
  LBX_LAUNCH_PDL(dp_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const unsigned int*)signal_pad_local, world,
                 epoch_dev, local_sync_dev, g_dp_spin_limit);
  return LBX_OK;
}

int lbx_dp_signal(void* const* signal_ptrs, int world, int rank, int slot_base, const unsigned int* epoch_dev,
                  unsigned int epoch_add, void* stream) {
  LBX_CHECK_ARG(signal_ptrs && epoch_dev, "NULL pointer argument");
  LBX_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world && slot_base >= 0, "bad rank / world / slot");
  LBX_LAUNCH_PDL(dp_signal_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (unsigned int* const*)signal_ptrs, world,
                 rank, slot_base, epoch_dev, epoch_add);
  return LBX_OK;
}

int lbx_dp_wait_slot(const void* signal_pad_local, int world, int slot_base, const unsigned int* epoch_dev,
                     unsigned int epoch_add, unsigned int* local_sync_dev, void* stream) {
  LBX_CHECK_ARG(signal_pad_local && epoch_dev && local_sync_dev, "NULL pointer argument");
  LBX_CHECK_ARG(world >= 1 && world <= 8 && slot_base >= 0, "bad world / slot");
  LBX_LAUNCH_PDL(dp_wait_slot_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (const unsigned int*)signal_pad_local,
                 world, slot_base, epoch_dev, epoch_add, local_sync_dev, g_dp_spin_limit);
  return LBX_OK;
}

int lbx_set_dp_blocks_per_sm(int n) {
  LBX_CHECK_ARG(n >= 1 && n <= 8, "blocks per SM must be in [1, 8]");
  g_dp_blocks_per_sm = n;
  return LBX_OK;
}

int lbx_set_dp_spin_limit(long long polls) {
  LBX_CHECK_ARG(polls >= 1, "spin limit must be >= 1");
  g_dp_spin_limit = polls;
  return LBX_OK;
}

}  // extern "C"
