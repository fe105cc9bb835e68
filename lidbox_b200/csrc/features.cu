// Feature kernels: frame + Hann + rFFT + |.|^p (+ sparse mel + log) for sm_100a.
//
// Replaces the TensorFlow op chain behind lidbox/features/audio.py:219-230 (spectrograms),
// :247-261 (linear_to_mel), lidbox/data/tf_utils.py:178 (log) and audio.py:167-174 (power_to_db).
//
// Fast path (fft_length == 512): one CTA owns a run of 32 consecutive frames of one utterance.  The samples the
// run covers are staged ONCE in shared memory (read amplification 1 + (L-step)/(32*step) instead of L/step), each
// half-warp computes one 512-point real FFT as a 256-point complex FFT (radix-16 in registers, one shared-memory
// transpose, radix-16 again) followed by a warp-shuffle split step, the power spectrum of the 32 frames is kept
// in shared memory and reduced against the band-compressed mel filterbank with lanes = frames (conflict-free),
// and the [32, n_mel] result leaves through shared memory as one contiguous, vectorised store.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace lbx {

// ------------------------------------------------------------------------------------------------------------
// small complex helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// forward 4-point DFT in place (W4 = -i)
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}

// position p of the in-place radix-16 output holds frequency KIDX(p)
__host__ __device__ constexpr int KIDX(int p) { return (p >> 2) + 4 * (p & 3); }
// inverse: frequency k lives at position PIDX(k)
__host__ __device__ constexpr int PIDX(int k) { return 4 * (k & 3) + (k >> 2); }

#define LBX_C1 0.92387953251128674f   // cos(pi/8)
#define LBX_S1 0.38268343236508977f   // sin(pi/8)
#define LBX_R2 0.70710678118654752f   // sqrt(1/2)

// forward 16-point DFT in place: input v[n] natural order, output v[p] = X[KIDX(p)]
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) fft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
  // v[4c + b] *= W16^{b c},  W16^e = (cos(pi e/8), -sin(pi e/8))
  v[5] = cmul(v[5], make_float2(LBX_C1, -LBX_S1));     // b=1,c=1 e=1
  v[6] = cmul(v[6], make_float2(LBX_R2, -LBX_R2));     // b=2,c=1 e=2
  v[7] = cmul(v[7], make_float2(LBX_S1, -LBX_C1));     // b=3,c=1 e=3
  v[9] = cmul(v[9], make_float2(LBX_R2, -LBX_R2));     // b=1,c=2 e=2
  v[10] = make_float2(v[10].y, -v[10].x);              // b=2,c=2 e=4 : * (-i)
  v[11] = cmul(v[11], make_float2(-LBX_R2, -LBX_R2));  // b=3,c=2 e=6
  v[13] = cmul(v[13], make_float2(LBX_S1, -LBX_C1));   // b=1,c=3 e=3
  v[14] = cmul(v[14], make_float2(-LBX_R2, -LBX_R2));  // b=2,c=3 e=6
  v[15] = cmul(v[15], make_float2(-LBX_C1, LBX_S1));   // b=3,c=3 e=9
#pragma unroll
  for (int c = 0; c < 4; ++c) fft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

// W32^{k} = (cos(pi k/16), -sin(pi k/16)), k = 0..15, as compile-time constants
__device__ __forceinline__ float2 w32(int k) {
  constexpr float c[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                           0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                           -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f,
                           -0.98078528040323043f};
  constexpr float s[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                           0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                           1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
  return make_float2(c[k], -s[k]);
}

__device__ __forceinline__ float apply_power(float mag2, int pw_mode, float power) {
  if (pw_mode == 2) return mag2;                 // |X|^2
  const float mag = sqrtf(mag2);
  if (pw_mode == 1) return mag;                  // |X|
  return powf(mag, power);                       // pow(abs(S), power), audio.py:230
}

// periodic Hann, tf.signal.hann_window(L, periodic=True): n = L + (1 - L%2) - 1
__device__ __forceinline__ float hann_value(int i, int L) {
  if (L == 1) return 1.0f;
  const float n = (float)(L + (1 - (L & 1)) - 1);
  return 0.5f - 0.5f * cosf(6.283185307179586f * (float)i / n);
}

// ------------------------------------------------------------------------------------------------------------
// fused 512-point kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int FR = 16;           // frames per CTA: 8 warps x 2 half-warps, one frame per half-warp, a single pass
constexpr int FUSED_THREADS = 256;
constexpr int P_STRIDE = 257;    // odd: the lanes = frames reads of the mel phase are conflict-free
constexpr int SCR_ROW = 17;      // float2 per transpose row (16 + 1 pad)
constexpr int SCR_FLOATS = 8 * 2 * 16 * SCR_ROW * 2;
constexpr int P_FLOATS = (FR * P_STRIDE + 3) & ~3;

struct FusedParams {
  const float* sig;
  const short* sig_i16;  // when non-NULL the input is 16-bit PCM and is converted as x / 32768 (what decode_wav does)
  float* out;
  long long N;
  long long T;
  int frame_length;
  int frame_step;
  int sig_smem;        // floats reserved for the staged signal run (multiple of 4, >= (FR-1)*step + 512 + 2)
  float power;
  // mel (MODE 1)
  int n_mel;
  const int* band_start;
  const int* band_len;
  const int* band_off;
  const float* band_w;
  int n_packed;
  int log_mode;
  float eps;
};

template <int PW>
__device__ __forceinline__ float power_of(float mag2, float power) {
  if (PW == 2) return mag2;                      // |X|^2
  if (PW == 1) return sqrtf(mag2);               // |X|
  return powf(sqrtf(mag2), power);               // pow(abs(S), power), audio.py:230
}

// MODE 0: power spectrogram [B,T,257]; MODE 1: (log-)mel [B,T,n_mel].  PW: 2 -> |X|^2, 1 -> |X|, 0 -> generic power.
template <int MODE, int PW>
__global__ void __launch_bounds__(FUSED_THREADS, 3) logmel512_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_sig = reinterpret_cast<float*>(smem_raw);
  float* s_win = s_sig + p.sig_smem;                         // 512 (zero beyond frame_length)
  float2* s_twA = reinterpret_cast<float2*>(s_win + 512);    // [q][l]: W256^{l * KIDX(q)}, 256 entries
  float2* s_w512 = s_twA + 256;                              // W512^k, k = 0..255
  float* s_P = reinterpret_cast<float*>(s_w512 + 256);       // FR * P_STRIDE
  float2* s_scr = reinterpret_cast<float2*>(s_P + P_FLOATS);
  float* s_out = reinterpret_cast<float*>(s_scr);            // aliases the transpose scratch (dead by then)
  float* s_bw = reinterpret_cast<float*>(s_scr) + SCR_FLOATS;
  int* s_bstart = reinterpret_cast<int*>(s_bw + ((p.n_packed + 3) & ~3));
  int* s_blen = s_bstart + p.n_mel;
  int* s_boff = s_blen + p.n_mel;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, half = lane >> 4, l16 = lane & 15;
  const int L = p.frame_length, step = p.frame_step;
  const long long t0 = (long long)blockIdx.x * FR;
  const int b = blockIdx.y;
  const int nf = (int)min((long long)FR, p.T - t0);

  // ---- tables that do not depend on earlier kernels: window and twiddles ----
  for (int i = tid; i < 512; i += FUSED_THREADS) {
    float w = 0.0f;
    if (i < L) {
      // periodic Hann, tf.signal.hann_window(L, periodic=True): n = L + (1 - L%2) - 1; cos(2 pi i / n) via cospi
      const float n = (float)(L + (1 - (L & 1)) - 1);
      w = L == 1 ? 1.0f : 0.5f - 0.5f * cospif(2.0f * (float)i / n);
    }
    s_win[i] = w;
  }
  {
    float sn, cs;
    sincospif((float)tid * (1.0f / 256.0f), &sn, &cs);       // W512^tid = exp(-2 pi i tid / 512)
    s_w512[tid] = make_float2(cs, -sn);
  }
  LBX_PDL_SYNC();
  // ---- stage the sample run and the mel tables ----
  {
    const long long s0 = t0 * step;
    const int n_valid = (nf - 1) * step + L;                  // samples this CTA actually needs (all in range)
    const float* g = p.sig + (long long)b * p.N + s0;
    if (p.sig_i16 != nullptr) {
      const short* gi = p.sig_i16 + (long long)b * p.N + s0;
      if ((reinterpret_cast<uintptr_t>(gi) & 7) == 0) {
        const int n4 = n_valid >> 2;                          // 4 samples per thread: 8-byte load, 16-byte store
        for (int i = tid; i < n4; i += FUSED_THREADS) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(gi) + i);
          reinterpret_cast<float4*>(s_sig)[i] =
              make_float4((float)(short)(u.x & 0xFFFFu) * (1.0f / 32768.0f), (float)(short)(u.x >> 16) * (1.0f / 32768.0f),
                          (float)(short)(u.y & 0xFFFFu) * (1.0f / 32768.0f), (float)(short)(u.y >> 16) * (1.0f / 32768.0f));
        }
        for (int i = (n4 << 2) + tid; i < n_valid; i += FUSED_THREADS) s_sig[i] = (float)__ldg(gi + i) * (1.0f / 32768.0f);
      } else {
        for (int i = tid; i < n_valid; i += FUSED_THREADS) s_sig[i] = (float)__ldg(gi + i) * (1.0f / 32768.0f);
      }
    } else if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const int n4 = n_valid >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(g);
      float4* s4 = reinterpret_cast<float4*>(s_sig);
      for (int i = tid; i < n4; i += FUSED_THREADS) s4[i] = __ldg(g4 + i);
      for (int i = (n4 << 2) + tid; i < n_valid; i += FUSED_THREADS) s_sig[i] = __ldg(g + i);
    } else {
      for (int i = tid; i < n_valid; i += FUSED_THREADS) s_sig[i] = __ldg(g + i);
    }
    for (int i = n_valid + tid; i < p.sig_smem; i += FUSED_THREADS) s_sig[i] = 0.0f;
    if (MODE == 1) {
      for (int i = tid; i < p.n_packed; i += FUSED_THREADS) s_bw[i] = __ldg(p.band_w + i);
      for (int i = tid; i < p.n_mel; i += FUSED_THREADS) {
        s_bstart[i] = __ldg(p.band_start + i);
        s_blen[i] = __ldg(p.band_len + i);
        s_boff[i] = __ldg(p.band_off + i);
      }
    }
  }
  __syncthreads();
  {
    // twiddles between the two radix-16 steps, stored so that a half-warp reads 16 consecutive entries:
    // s_twA[q*16 + l] = W256^{l * KIDX(q)};  W256^e = W512^{2e} (e < 128) = -W512^{2e-256} (e >= 128)
    const int q = tid >> 4, l = tid & 15;
    const int e = (l * KIDX(q)) & 255;
    const float2 w = e < 128 ? s_w512[2 * e] : s_w512[2 * e - 256];
    s_twA[tid] = e < 128 ? w : make_float2(-w.x, -w.y);
  }
  __syncthreads();

  const int f = half * 8 + warp;                              // frame within the CTA's run
  const float* fs = s_sig + f * step;
  float2* my_scr = s_scr + (warp * 2 + half) * 16 * SCR_ROW;
  float2 v[16];
  // z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1],  n = 16 n1 + l16 (rows beyond the window multiply by zero)
  if ((step & 1) == 0) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float2 x = *reinterpret_cast<const float2*>(fs + 32 * n1 + 2 * l16);
      const float2 w = *reinterpret_cast<const float2*>(s_win + 32 * n1 + 2 * l16);
      v[n1] = make_float2(x.x * w.x, x.y * w.y);
    }
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int i = 32 * n1 + 2 * l16;
      v[n1] = make_float2(fs[i] * s_win[i], fs[i + 1] * s_win[i + 1]);
    }
  }
  fft16(v);
#pragma unroll
  for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], s_twA[q * 16 + l16]);
#pragma unroll
  for (int q = 0; q < 16; ++q) my_scr[KIDX(q) * SCR_ROW + l16] = v[q];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = my_scr[l16 * SCR_ROW + n2];
  fft16(v);                                                   // v[q] = Z[l16 + 16 KIDX(q)]

  // split step.  With E = (Z[k] + conj Z[256-k]) / 2, O = (Z[k] - conj Z[256-k]) / 2i and t = W512^k O:
  //   X[k] = E + t,  X[256-k] = conj(E - t)  ->  both power bins from one evaluation; only k2 = KIDX(q) < 8 is walked.
  // The partner Z[256-k] lives in lane (16 - l16) & 15, register 15 - q (lane 0 pairs with itself).
  float* Prow = s_P + f * P_STRIDE;
  const int src_lane = (half << 4) | ((16 - l16) & 15);
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int q = (qi >> 1) * 4 + (qi & 1);                   // 0,1,4,5,8,9,12,13
    const int k2 = KIDX(q);
    float cx = __shfl_sync(0xffffffffu, v[15 - q].x, src_lane);
    float cy = __shfl_sync(0xffffffffu, v[15 - q].y, src_lane);
    if (l16 == 0) {                                           // k1 == 0: partner 16*((16-k2)&15) is in this lane
      cx = v[PIDX((16 - k2) & 15)].x;
      cy = v[PIDX((16 - k2) & 15)].y;
    }
    const float a = v[q].x, bb = v[q].y;
    const float ex = a + cx, ey = bb - cy;                    // 2 E
    const float2 t = cmul(s_w512[l16 + 16 * k2], make_float2(bb + cy, cx - a));   // W512^k * 2 O
    const float x1 = ex + t.x, y1 = ey + t.y, x2 = ex - t.x, y2 = ey - t.y;
    const int k = l16 + 16 * k2;
    Prow[k] = power_of<PW>(0.25f * fmaf(x1, x1, y1 * y1), p.power);
    Prow[256 - k] = power_of<PW>(0.25f * fmaf(x2, x2, y2 * y2), p.power);
  }
  if (l16 == 0) Prow[128] = power_of<PW>(fmaf(v[PIDX(8)].x, v[PIDX(8)].x, v[PIDX(8)].y * v[PIDX(8)].y), p.power);
  __syncthreads();

  if (MODE == 0) {
    float* dst = p.out + ((long long)b * p.T + t0) * 257;
    const int total = nf * 257;                               // s_P rows are contiguous (stride 257)
    for (int i = tid; i < total; i += FUSED_THREADS) dst[i] = s_P[i];
  } else {
    const int n_mel = p.n_mel;
    const int mf = tid & 15;                                  // frame
    const float* Pf = s_P + mf * P_STRIDE;
    for (int m = tid >> 4; m < n_mel; m += 16) {
      const int start = s_bstart[m], len = s_blen[m];
      const float* w = s_bw + s_boff[m];
      const float* Pk = Pf + start;
      float acc = 0.0f;
      for (int j = 0; j < len; ++j) acc = fmaf(Pk[j], w[j], acc);
      if (p.log_mode == 1) acc = logf(acc + p.eps);
      s_out[mf * n_mel + m] = acc;
    }
    __syncthreads();
    float* dst = p.out + ((long long)b * p.T + t0) * n_mel;
    const int total = nf * n_mel;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const int n4 = total >> 2;
      for (int i = tid; i < n4; i += FUSED_THREADS)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_out)[i];
      for (int i = (n4 << 2) + tid; i < total; i += FUSED_THREADS) dst[i] = s_out[i];
    } else {
      for (int i = tid; i < total; i += FUSED_THREADS) dst[i] = s_out[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// generic power-of-two path (any fft_length in [32, 4096]): one CTA per frame, shared-memory radix-2
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stft_generic_kernel(const float* __restrict__ sig, float* __restrict__ out,
                                                          long long N, long long T, int L, int step, int nfft,
                                                          int log2n, int pw_mode, float power) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  const long long t = blockIdx.x;
  const int b = blockIdx.y;
  const float* x = sig + (long long)b * N + t * step;
  for (int i = threadIdx.x; i < nfft; i += blockDim.x) {
    const int j = (int)(__brev((unsigned)i) >> (32 - log2n));
    const float val = (i < L) ? __ldg(x + i) * hann_value(i, L) : 0.0f;
    buf[j] = make_float2(val, 0.0f);
  }
  __syncthreads();
  for (int s = 1; s <= log2n; ++s) {
    const int m = 1 << s, hm = m >> 1;
    for (int idx = threadIdx.x; idx < (nfft >> 1); idx += blockDim.x) {
      const int grp = idx >> (s - 1), j = idx & (hm - 1);
      float sn, cs;
      sincospif((float)(2 * j) / (float)m, &sn, &cs);
      const float2 w = make_float2(cs, -sn);
      const int i0 = grp * m + j;
      const float2 a = buf[i0], bb = cmul(buf[i0 + hm], w);
      buf[i0] = cadd(a, bb);
      buf[i0 + hm] = csub(a, bb);
    }
    __syncthreads();
  }
  const int K = (nfft >> 1) + 1;
  float* dst = out + ((long long)b * T + t) * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float2 z = buf[k];
    dst[k] = apply_power(fmaf(z.x, z.x, z.y * z.y), pw_mode, power);
  }
}

// standalone band-compressed mel projection: S [rows, n_bins] -> out [rows, n_mel]
constexpr int MEL_ROWS = 4;
__global__ void __launch_bounds__(256) linear_to_mel_kernel(const float* __restrict__ S, float* __restrict__ out,
                                                           long long rows, int n_bins, int n_mel,
                                                           const int* __restrict__ band_start,
                                                           const int* __restrict__ band_len,
                                                           const int* __restrict__ band_off,
                                                           const float* __restrict__ band_w, int log_mode, float eps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  const long long r0 = (long long)blockIdx.x * MEL_ROWS;
  const int nr = (int)min((long long)MEL_ROWS, rows - r0);
  const float* src = S + r0 * n_bins;
  for (int i = threadIdx.x; i < nr * n_bins; i += blockDim.x) tile[i] = __ldg(src + i);
  __syncthreads();
  float* dst = out + r0 * n_mel;
  for (int i = threadIdx.x; i < nr * n_mel; i += blockDim.x) {
    const int r = i / n_mel, m = i - r * n_mel;
    const int start = __ldg(band_start + m), len = __ldg(band_len + m);
    const float* w = band_w + __ldg(band_off + m);
    const float* row = tile + r * n_bins + start;
    float acc = 0.0f;
    for (int j = 0; j < len; ++j) acc = fmaf(row[j], __ldg(w + j), acc);
    if (log_mode == 1) acc = logf(acc + eps);
    dst[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// power_to_db: global max of max(amin, S) then the elementwise map
// ------------------------------------------------------------------------------------------------------------
__global__ void ptdb_init_kernel(float* ws, float amin) { ws[0] = amin; }

__global__ void __launch_bounds__(256) ptdb_max_kernel(const float* __restrict__ S, long long n, float amin, float* ws) {
  float m = amin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, __ldg(S + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
    // values are >= amin > 0 (or amin itself): positive floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(ws), __float_as_int(m));
  }
}

__global__ void __launch_bounds__(256) ptdb_map_kernel(const float* __restrict__ S, float* __restrict__ out, long long n,
                                                      float amin, float top_db, const float* __restrict__ ws) {
  const float ln10 = logf(10.0f);
  const float ref = logf(ws[0]) / ln10;                    // log10(max(amin, max_all S)), audio.py:163-164,173
  // max_all(db) is attained by the max element: 20*(ref - ref) = 0, so the floor is -top_db (audio.py:174)
  const float floor_db = 20.0f * (ref - ref) - top_db;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float db = 20.0f * (logf(fmaxf(amin, __ldg(S + i))) / ln10 - ref);
    out[i] = fmaxf(db, floor_db);
  }
}

__global__ void __launch_bounds__(256) check_finite_kernel(const float* __restrict__ x, long long n, int* flag) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    bad |= !isfinite(v);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

// ------------------------------------------------------------------------------------------------------------
// host-side launch logic
// ------------------------------------------------------------------------------------------------------------
static int pw_mode_of(float power) { return power == 2.0f ? 2 : (power == 1.0f ? 1 : 0); }

static int fused_sig_smem(int L, int step) {
  (void)L;
  long long n = (long long)(FR - 1) * step + 512 + 2;       // every frame reads all 512 points (window is 0 past L)
  return (int)((n + 3) & ~3LL);
}

static size_t fused_smem_bytes(int L, int step, int n_mel, int n_packed) {
  size_t floats = (size_t)fused_sig_smem(L, step) + 512 + 512 /* s_twA */ + 512 /* s_w512 */ + (size_t)P_FLOATS +
                  SCR_FLOATS + (size_t)((n_packed + 3) & ~3) + 3 * (size_t)n_mel;
  return floats * 4;
}

static bool fused_ok(int L, int step, int nfft, int n_mel, int n_packed) {
  if (nfft != 512 || L > 512 || L < 1 || step < 1) return false;
  if (n_mel > 256 || FR * n_mel > SCR_FLOATS) return false;
  return fused_smem_bytes(L, step, n_mel, n_packed) <= 110 * 1024;
}

static int check_stft_args(const float* sig, long long B, long long N, int L, int step, int nfft) {
  LBX_CHECK_ARG(B >= 0 && N >= 0, "negative shape B=%lld N=%lld", B, N);
  LBX_CHECK_ARG(B <= 65535, "batch %lld exceeds the grid limit 65535; split the batch", B);
  LBX_CHECK_ARG(L >= 1 && step >= 1, "frame_length=%d and frame_step=%d must be >= 1", L, step);
  if (!is_pow2(nfft) || nfft < 32 || nfft > 4096)
    return set_error(LBX_EUNSUPPORTED, "fft_length=%d: only powers of two in [32, 4096] are implemented", nfft);
  if (nfft < L)
    return set_error(LBX_EUNSUPPORTED, "fft_length=%d < frame_length=%d (cropping frames) is not implemented", nfft, L);
  LBX_CHECK_ARG(sig != nullptr || B * N == 0, "sig is NULL");
  return LBX_OK;
}

template <int MODE, int PW>
static int launch_fused_pw(const FusedParams& p, long long B, size_t smem, cudaStream_t st) {
  LBX_CUDA(cudaFuncSetAttribute(logmel512_kernel<MODE, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(p.T, FR), (unsigned)B);
  LBX_LAUNCH_PDL((logmel512_kernel<MODE, PW>), grid, dim3(FUSED_THREADS), smem, st, p);
  return LBX_OK;
}

template <int MODE>
static int launch_fused(const FusedParams& p, long long B, size_t smem, cudaStream_t st) {
  const int pw = pw_mode_of(p.power);
  if (pw == 2) return launch_fused_pw<MODE, 2>(p, B, smem, st);
  if (pw == 1) return launch_fused_pw<MODE, 1>(p, B, smem, st);
  return launch_fused_pw<MODE, 0>(p, B, smem, st);
}

static int launch_generic_stft(const float* sig, long long B, long long N, long long T, int L, int step, int nfft,
                               float power, float* out, cudaStream_t st) {
  int log2n = 0;
  while ((1 << log2n) < nfft) ++log2n;
  dim3 grid((unsigned)T, (unsigned)B);
  LBX_CHECK_ARG(T <= 2147483647LL, "too many frames");
  stft_generic_kernel<<<grid, 128, (size_t)nfft * sizeof(float2), st>>>(sig, out, N, T, L, step, nfft, log2n,
                                                                         pw_mode_of(power), power);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

}  // namespace lbx

using namespace lbx;

extern "C" {

int lbx_spectrogram_f32(const float* sig, long long B, long long N, int frame_length, int frame_step, int fft_length,
                        float power, float* out, void* stream) {
  int rc = check_stft_args(sig, B, N, frame_length, frame_step, fft_length);
  if (rc) return rc;
  const long long T = lbx_num_frames(N, frame_length, frame_step);
  if (B == 0 || T == 0) return LBX_OK;
  LBX_CHECK_ARG(out != nullptr, "out is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_ok(frame_length, frame_step, fft_length, 0, 0)) {
    FusedParams p{};
    p.sig = sig; p.out = out; p.N = N; p.T = T;
    p.frame_length = frame_length; p.frame_step = frame_step;
    p.sig_smem = fused_sig_smem(frame_length, frame_step);
    p.power = power;
    return launch_fused<0>(p, B, fused_smem_bytes(frame_length, frame_step, 0, 0), st);
  }
  return launch_generic_stft(sig, B, N, T, frame_length, frame_step, fft_length, power, out, st);
}

int lbx_linear_to_mel_f32(const float* S, long long rows, int n_bins, int n_mel, const int* band_start,
                          const int* band_len, const int* band_off, const float* band_w, int n_packed, int log_mode,
                          float eps, float* out, void* stream) {
  LBX_CHECK_ARG(rows >= 0 && n_bins >= 1 && n_mel >= 1, "bad shape rows=%lld n_bins=%d n_mel=%d", rows, n_bins, n_mel);
  LBX_CHECK_ARG(log_mode == 0 || log_mode == 1, "log_mode must be 0 or 1");
  (void)n_packed;
  if (rows == 0) return LBX_OK;
  LBX_CHECK_ARG(S && out && band_start && band_len && band_off && band_w, "NULL pointer argument");
  const size_t smem = (size_t)MEL_ROWS * n_bins * sizeof(float);
  if (smem > 200 * 1024) return set_error(LBX_EUNSUPPORTED, "n_bins=%d too large", n_bins);
  if (smem > 48 * 1024)
    LBX_CUDA(cudaFuncSetAttribute(linear_to_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = ceil_div(rows, MEL_ROWS);
  LBX_CHECK_ARG(blocks <= 2147483647LL, "too many rows");
  linear_to_mel_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(S, out, rows, n_bins, n_mel, band_start,
                                                                             band_len, band_off, band_w, log_mode, eps);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

size_t lbx_logmel_workspace_bytes(long long B, long long N, int frame_length, int frame_step, int fft_length,
                                  int n_mel) {
  if (frame_length < 1 || frame_step < 1 || fft_length < 2) return 0;
  // n_packed is not known here; the fused kernel's table area is bounded by n_bins * 2 weights for triangular banks
  if (fused_ok(frame_length, frame_step, fft_length, n_mel, 2 * (fft_length / 2 + 1))) return 0;
  const long long T = lbx_num_frames(N, frame_length, frame_step);
  return (size_t)B * (size_t)T * (size_t)(fft_length / 2 + 1) * sizeof(float);
}

int lbx_logmel_f32(const float* sig, long long B, long long N, int frame_length, int frame_step, int fft_length,
                   float power, int n_mel, const int* band_start, const int* band_len, const int* band_off,
                   const float* band_w, int n_packed, int log_mode, float eps, float* out, void* workspace,
                   size_t workspace_bytes, void* stream) {
  int rc = check_stft_args(sig, B, N, frame_length, frame_step, fft_length);
  if (rc) return rc;
  LBX_CHECK_ARG(n_mel >= 1 && n_packed >= 0, "bad n_mel=%d n_packed=%d", n_mel, n_packed);
  LBX_CHECK_ARG(log_mode == 0 || log_mode == 1, "log_mode must be 0 or 1");
  const long long T = lbx_num_frames(N, frame_length, frame_step);
  if (B == 0 || T == 0) return LBX_OK;
  LBX_CHECK_ARG(out && band_start && band_len && band_off && band_w, "NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_ok(frame_length, frame_step, fft_length, n_mel, n_packed)) {
    FusedParams p{};
    p.sig = sig; p.out = out; p.N = N; p.T = T;
    p.frame_length = frame_length; p.frame_step = frame_step;
    p.sig_smem = fused_sig_smem(frame_length, frame_step);
    p.power = power;
    p.n_mel = n_mel; p.band_start = band_start; p.band_len = band_len; p.band_off = band_off; p.band_w = band_w;
    p.n_packed = n_packed; p.log_mode = log_mode; p.eps = eps;
    return launch_fused<1>(p, B, fused_smem_bytes(frame_length, frame_step, n_mel, n_packed), st);
  }
  const int K = fft_length / 2 + 1;
  const size_t need = (size_t)B * (size_t)T * (size_t)K * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need)
    return set_error(LBX_EWORKSPACE, "logmel needs a %zu-byte workspace for this configuration (got %zu)", need,
                     workspace_bytes);
  rc = launch_generic_stft(sig, B, N, T, frame_length, frame_step, fft_length, power, (float*)workspace, st);
  if (rc) return rc;
  return lbx_linear_to_mel_f32((const float*)workspace, B * T, K, n_mel, band_start, band_len, band_off, band_w,
                               n_packed, log_mode, eps, out, stream);
}

int lbx_logmel_i16(const short* pcm, long long B, long long N, int frame_length, int frame_step, int fft_length,
                   float power, int n_mel, const int* band_start, const int* band_len, const int* band_off,
                   const float* band_w, int n_packed, int log_mode, float eps, float* out, void* stream) {
  int rc = check_stft_args(reinterpret_cast<const float*>(pcm), B, N, frame_length, frame_step, fft_length);
  if (rc) return rc;
  LBX_CHECK_ARG(n_mel >= 1 && n_packed >= 0, "bad n_mel=%d n_packed=%d", n_mel, n_packed);
  LBX_CHECK_ARG(log_mode == 0 || log_mode == 1, "log_mode must be 0 or 1");
  const long long T = lbx_num_frames(N, frame_length, frame_step);
  if (B == 0 || T == 0) return LBX_OK;
  LBX_CHECK_ARG(out && band_start && band_len && band_off && band_w, "NULL pointer argument");
  if (!fused_ok(frame_length, frame_step, fft_length, n_mel, n_packed))
    return set_error(LBX_EUNSUPPORTED, "16-bit PCM input is served by the fused 512-point configuration only");
  FusedParams p{};
  p.sig = nullptr; p.sig_i16 = pcm; p.out = out; p.N = N; p.T = T;
  p.frame_length = frame_length; p.frame_step = frame_step;
  p.sig_smem = fused_sig_smem(frame_length, frame_step);
  p.power = power;
  p.n_mel = n_mel; p.band_start = band_start; p.band_len = band_len; p.band_off = band_off; p.band_w = band_w;
  p.n_packed = n_packed; p.log_mode = log_mode; p.eps = eps;
  return launch_fused<1>(p, B, fused_smem_bytes(frame_length, frame_step, n_mel, n_packed), (cudaStream_t)stream);
}

int lbx_power_to_db_f32(const float* S, long long numel, float amin, float top_db, float* out, void* workspace,
                        void* stream) {
  LBX_CHECK_ARG(numel >= 0, "negative numel");
  if (numel == 0) return LBX_OK;
  LBX_CHECK_ARG(S && out && workspace, "NULL pointer argument");
  LBX_CHECK_ARG(amin > 0.0f, "amin must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const int blocks = (int)min((long long)148 * 8, ceil_div(numel, 256));
  ptdb_init_kernel<<<1, 1, 0, st>>>(ws, amin);
  LBX_LAUNCH_CHECK();
  ptdb_max_kernel<<<blocks, 256, 0, st>>>(S, numel, amin, ws);
  LBX_LAUNCH_CHECK();
  ptdb_map_kernel<<<blocks, 256, 0, st>>>(S, out, numel, amin, top_db, ws);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

int lbx_check_finite_f32(const float* x, long long numel, int* flag_dev, void* stream) {
  LBX_CHECK_ARG(numel >= 0, "negative numel");
  if (numel == 0) return LBX_OK;
  LBX_CHECK_ARG(x && flag_dev, "NULL pointer argument");
  const int blocks = (int)min((long long)148 * 8, ceil_div(numel, 256));
  check_finite_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, numel, flag_dev);
  LBX_LAUNCH_CHECK();
  return LBX_OK;
}

int lbx_logmel_f32_host(const float* sig_host, long long B, long long N, int sample_rate, int frame_length_ms,
                        int frame_step_ms, int fft_length, float power, int n_mel, float fmin, float fmax,
                        int log_mode, float eps, float* out_host, float* dev_sig, float* dev_out, void* dev_tables,
                        size_t dev_tables_bytes, void* stream) {
  const int L = lbx_ms_to_frames(sample_rate, frame_length_ms);
  const int step = lbx_ms_to_frames(sample_rate, frame_step_ms);
  int rc = check_stft_args(dev_sig, B, N, L, step, fft_length);
  if (rc) return rc;
  const int K = fft_length / 2 + 1;
  const long long T = lbx_num_frames(N, L, step);
  if (B == 0 || T == 0) return LBX_OK;
  LBX_CHECK_ARG(sig_host && out_host && dev_sig && dev_out && dev_tables, "NULL pointer argument");
  if (lbx_logmel_workspace_bytes(B, N, L, step, fft_length, n_mel) != 0)
    return set_error(LBX_EUNSUPPORTED, "lbx_logmel_f32_host only serves the fused 512-point configuration");
  std::vector<float> W((size_t)K * n_mel), packed((size_t)K * n_mel);
  std::vector<int> start(n_mel), len(n_mel), off(n_mel);
  rc = lbx_mel_weight_matrix(n_mel, K, sample_rate, fmin, fmax, W.data());
  if (rc) return rc;
  const int n_packed = lbx_mel_pack_bands(W.data(), K, n_mel, start.data(), len.data(), off.data(), packed.data());
  if (n_packed < 0) return n_packed;
  const size_t need = (size_t)(3 * n_mel) * sizeof(int) + (size_t)n_packed * sizeof(float);
  if (dev_tables_bytes < need) return set_error(LBX_EWORKSPACE, "dev_tables needs %zu bytes", need);
  cudaStream_t st = (cudaStream_t)stream;
  int* d_start = (int*)dev_tables;
  int* d_len = d_start + n_mel;
  int* d_off = d_len + n_mel;
  float* d_w = (float*)(d_off + n_mel);
  LBX_CUDA(cudaMemcpyAsync(d_start, start.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice, st));
  LBX_CUDA(cudaMemcpyAsync(d_len, len.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice, st));
  LBX_CUDA(cudaMemcpyAsync(d_off, off.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice, st));
  LBX_CUDA(cudaMemcpyAsync(d_w, packed.data(), (size_t)n_packed * sizeof(float), cudaMemcpyHostToDevice, st));
  LBX_CUDA(cudaMemcpyAsync(dev_sig, sig_host, (size_t)B * N * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = lbx_logmel_f32(dev_sig, B, N, L, step, fft_length, power, n_mel, d_start, d_len, d_off, d_w, n_packed, log_mode,
                      eps, dev_out, nullptr, 0, stream);
  if (rc) return rc;
  LBX_CUDA(cudaMemcpyAsync(out_host, dev_out, (size_t)B * T * n_mel * sizeof(float), cudaMemcpyDeviceToHost, st));
  LBX_CUDA(cudaStreamSynchronize(st));   // the host vectors above must outlive the async copies
  return LBX_OK;
}

}  // extern "C"
