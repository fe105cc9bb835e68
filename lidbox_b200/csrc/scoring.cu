// Consumer side of the path: fixed-length signal chunks in front of the feature stage,
// chunk-prediction averaging behind the model, and the average detection cost C_avg on device.
//   lidbox/data/steps.py:579-632  create_signal_chunks        -> signal_chunks_kernel
//   lidbox/util.py:41-57          merge_chunk_predictions     -> group_mean_kernel
//   lidbox/metrics.py:6-119       AverageDetectionCost        -> cavg_update_kernel, cavg_result_kernel
#include "common.cuh"
#include <math.h>

namespace lbx {

// out[(b*C + c), i] = sig[b, c*step + i], zero where c*step + i >= N (the padded tail of the last chunk)
__global__ void __launch_bounds__(256) signal_chunks_kernel(const float* __restrict__ sig, long long N, long long C,
                                                           long long L, long long step, long long total4,
                                                           float* __restrict__ out) {
  LBX_PDL_SYNC();
  const long long L4 = (L + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4;
       q += (long long)gridDim.x * blockDim.x) {
    const long long row = q / L4, i = (q - row * L4) * 4;
    const long long b = row / C, c = row - b * C;
    const float* src = sig + b * N + c * step;
    float* dst = out + row * L;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i + j < L) dst[i + j] = (c * step + i + j < N) ? src[i + j] : 0.0f;
  }
}

// out[g, :] = mean over rows row_index[group_offsets[g] .. group_offsets[g+1]) of pred[row, :]
__global__ void __launch_bounds__(128) group_mean_kernel(const float* __restrict__ pred, const long long* __restrict__ row_index,
                                                        const long long* __restrict__ group_offsets, int D,
                                                        float* __restrict__ out) {
  LBX_PDL_SYNC();
  const long long g = blockIdx.x;
  const long long lo = group_offsets[g], hi = group_offsets[g + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.0f;
    for (long long r = lo; r < hi; ++r) s += pred[row_index[r] * D + d];
    out[g * D + d] = hi > lo ? s / (float)(hi - lo) : 0.0f;
  }
}

// One block = one class column n x one slab of samples.  Counters of the slab are accumulated in shared memory
// ([2 + 2N] x Th floats) and flushed with one global atomic per nonzero counter; when they do not fit into shared
// memory the block adds to the global counters directly.
template <bool SMEM>
__global__ void __launch_bounds__(256) cavg_update_kernel(const float* __restrict__ onehot, const int* __restrict__ labels,
                                                         const float* __restrict__ pred, long long B, int N,
                                                         const float* __restrict__ thr, int Th, int slab,
                                                         float* __restrict__ tp, float* __restrict__ fn,
                                                         float* __restrict__ fp_pairs, float* __restrict__ tn_pairs) {
  LBX_PDL_SYNC();
  extern __shared__ float s_cnt[];      // SMEM: tp[Th] fn[Th] fp[N][Th] tn[N][Th]; then int label[slab]
  const int n = blockIdx.x;
  const long long b0 = (long long)blockIdx.y * slab;
  const int nb = (int)min((long long)slab, B - b0);
  const int n_cnt = SMEM ? (2 + 2 * N) * Th : 0;
  int* s_label = reinterpret_cast<int*>(s_cnt + n_cnt);
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) s_cnt[i] = 0.0f;
  // metrics.py:56: label index = argmax of the dense row (first maximum); sparse labels: one_hot() semantics, an
  // out-of-range label gives an all-zero row whose argmax is 0
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    int l = 0;
    if (onehot != nullptr) {
      const float* row = onehot + (b0 + i) * N;
      float best = row[0];
      for (int m = 1; m < N; ++m)
        if (row[m] > best) { best = row[m]; l = m; }
    } else {
      l = labels[b0 + i];
      if (l < 0 || l >= N) l = -1;
    }
    s_label[i] = l;
  }
  __syncthreads();
  float* c_tp = SMEM ? s_cnt : tp + (long long)n * Th;
  float* c_fn = SMEM ? s_cnt + Th : fn + (long long)n * Th;
  for (int idx = threadIdx.x; idx < nb * Th; idx += blockDim.x) {
    const int i = idx / Th, t = idx - i * Th;
    const int l = s_label[i];
    const float p = pred[(b0 + i) * N + n];
    const float w = onehot != nullptr ? onehot[(b0 + i) * N + n] : (l == n ? 1.0f : 0.0f);
    const float th = thr[t];
    const bool pos = p >= th, neg = p < th;            // both false for a NaN score (metrics.py:60-61)
    if (w != 0.0f) {
      if (pos) atomicAdd(c_tp + t, w);
      if (neg) atomicAdd(c_fn + t, w);
    } else {
      const int lrow = l < 0 ? 0 : l;
      float* c_fp = SMEM ? s_cnt + (2 + lrow) * Th : fp_pairs + ((long long)lrow * N + n) * Th;
      float* c_tn = SMEM ? s_cnt + (2 + N + lrow) * Th : tn_pairs + ((long long)lrow * N + n) * Th;
      if (pos) atomicAdd(c_fp + t, 1.0f);
      if (neg) atomicAdd(c_tn + t, 1.0f);
    }
  }
  if (!SMEM) return;
  __syncthreads();
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) {
    const float v = s_cnt[i];
    if (v == 0.0f) continue;
    const int r = i / Th, t = i - r * Th;
    if (r == 0) atomicAdd(tp + (long long)n * Th + t, v);
    else if (r == 1) atomicAdd(fn + (long long)n * Th + t, v);
    else if (r < 2 + N) atomicAdd(fp_pairs + ((long long)(r - 2) * N + n) * Th + t, v);
    else atomicAdd(tn_pairs + ((long long)(r - 2 - N) * N + n) * Th + t, v);
  }
}

__device__ __forceinline__ float divide_no_nan(float a, float b) { return b == 0.0f ? 0.0f : a / b; }

// metrics.py:74-99: one thread per threshold, then a block-wide minimum
__global__ void __launch_bounds__(256) cavg_result_kernel(const float* __restrict__ tp, const float* __restrict__ fn,
                                                         const float* __restrict__ fp_pairs,
                                                         const float* __restrict__ tn_pairs, int N, int Th, float C_miss,
                                                         float C_fa, float P_tar, float* __restrict__ cavg,
                                                         float* __restrict__ cavg_min) {
  LBX_PDL_SYNC();
  __shared__ float red[256];
  float best = INFINITY;
  for (int t = threadIdx.x; t < Th; t += blockDim.x) {
    float miss = 0.0f;
    for (int n = 0; n < N; ++n) {
      const float a = fn[(long long)n * Th + t], b = tp[(long long)n * Th + t];
      miss += divide_no_nan(a, a + b);
    }
    miss /= (float)N;
    float fa = 0.0f;
    for (int l = 0; l < N; ++l) {
      float s = 0.0f;
      for (int m = 0; m < N; ++m) {
        const long long i = ((long long)l * N + m) * Th + t;
        s += divide_no_nan(fp_pairs[i], fp_pairs[i] + tn_pairs[i]);
      }
      fa += divide_no_nan(s, (float)(N - 1));
    }
    fa /= (float)N;
    const float c = C_miss * P_tar * miss + C_fa * (1.0f - P_tar) * fa;
    if (cavg != nullptr) cavg[t] = c;
    best = fminf(best, c);
  }
  red[threadIdx.x] = best;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *cavg_min = red[0];
}

}  // namespace lbx

using namespace lbx;

extern "C" {

long long lbx_num_signal_chunks(long long N, long long chunk_length, long long chunk_step, long long max_pad) {
  if (chunk_length <= 0 || chunk_step <= 0 || N < 0) return -1;
  // steps.py:607-613: floor division as tf's int32 `//`
  long long d = N - chunk_length;
  long long q = d / chunk_step;
  if ((d % chunk_step != 0) && (d < 0)) --q;
  long long full = 1 + q;
  if (full < 0) full = 0;
  const long long last = N - full * chunk_step;
  if (last < chunk_length && chunk_length <= last + max_pad) return full + 1;   // one zero-padded chunk more
  return full;
}

int lbx_signal_chunks_f32(const float* sig, long long B, long long N, long long chunk_length, long long chunk_step,
                          long long num_chunks, float* out, void* stream) {
  LBX_CHECK_ARG(B >= 0 && N >= 0 && chunk_length > 0 && chunk_step > 0 && num_chunks >= 0, "bad chunk geometry");
  if (B == 0 || num_chunks == 0) return LBX_OK;
  LBX_CHECK_ARG(sig != nullptr && out != nullptr, "null pointer");
  LBX_CHECK_ARG((num_chunks - 1) * chunk_step < N || N == 0, "chunk %lld starts behind the end of the signal",
                num_chunks - 1);
  const long long total4 = B * num_chunks * ((chunk_length + 3) / 4);
  const long long blocks = ceil_div(total4, 256);
  const unsigned grid = (unsigned)(blocks < 148LL * 16 ? blocks : 148LL * 16);
  LBX_LAUNCH_PDL(signal_chunks_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, sig, N, num_chunks,
                 chunk_length, chunk_step, total4, out);
  return LBX_OK;
}

int lbx_group_mean_f32(const float* pred, const long long* row_index, const long long* group_offsets,
                       long long num_groups, int D, float* out, void* stream) {
  LBX_CHECK_ARG(num_groups >= 0 && D >= 1, "bad shape");
  if (num_groups == 0) return LBX_OK;
  LBX_CHECK_ARG(pred != nullptr && row_index != nullptr && group_offsets != nullptr && out != nullptr,
                "null pointer");
  LBX_CHECK_ARG(num_groups <= 0x7fffffffLL, "too many groups");
  LBX_LAUNCH_PDL(group_mean_kernel, dim3((unsigned)num_groups), dim3(128), 0, (cudaStream_t)stream, pred, row_index,
                 group_offsets, D, out);
  return LBX_OK;
}

int lbx_cavg_update_f32(const float* onehot, const int* labels, const float* pred, long long B, int N,
                        const float* thresholds, int num_thresholds, float* tp, float* fn, float* fp_pairs,
                        float* tn_pairs, void* stream) {
  LBX_CHECK_ARG(N >= 2, "C_avg is undefined for less than 2 classes");
  LBX_CHECK_ARG(num_thresholds >= 1 && B >= 0, "bad shape");
  LBX_CHECK_ARG((onehot != nullptr) != (labels != nullptr), "pass either dense one-hot rows or sparse labels");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(pred && thresholds && tp && fn && fp_pairs && tn_pairs, "null pointer");
  const int slab = 256;
  const size_t cnt_bytes = (size_t)(2 + 2 * N) * num_thresholds * sizeof(float);
  const size_t lab_bytes = slab * sizeof(int);
  const long long slabs = ceil_div(B, slab);
  LBX_CHECK_ARG(slabs <= 65535, "batch too large for one update (max %d samples)", 65535 * slab);
  dim3 grid((unsigned)N, (unsigned)slabs);
  if (cnt_bytes + lab_bytes <= 200 * 1024) {
    if (cnt_bytes + lab_bytes > 48 * 1024)
      LBX_CUDA(cudaFuncSetAttribute(cavg_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(cnt_bytes + lab_bytes)));
    LBX_LAUNCH_PDL(cavg_update_kernel<true>, grid, dim3(256), cnt_bytes + lab_bytes, (cudaStream_t)stream, onehot,
                   labels, pred, B, N, thresholds, num_thresholds, slab, tp, fn, fp_pairs, tn_pairs);
  } else {
    LBX_LAUNCH_PDL(cavg_update_kernel<false>, grid, dim3(256), lab_bytes, (cudaStream_t)stream, onehot, labels, pred,
                   B, N, thresholds, num_thresholds, slab, tp, fn, fp_pairs, tn_pairs);
  }
  return LBX_OK;
}

int lbx_cavg_result_f32(const float* tp, const float* fn, const float* fp_pairs, const float* tn_pairs, int N,
                        int num_thresholds, float C_miss, float C_fa, float P_tar, float* cavg_per_threshold,
                        float* cavg_min, void* stream) {
  LBX_CHECK_ARG(N >= 2 && num_thresholds >= 1, "bad shape");
  LBX_CHECK_ARG(tp && fn && fp_pairs && tn_pairs && cavg_min, "null pointer");
  LBX_LAUNCH_PDL(cavg_result_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, tp, fn, fp_pairs, tn_pairs, N,
                 num_thresholds, C_miss, C_fa, P_tar, cavg_per_threshold, cavg_min);
  return LBX_OK;
}

}  // extern "C"
