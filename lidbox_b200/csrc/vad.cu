// Energy VAD in front of the feature stage, batched over utterances:
// lidbox/features/audio.py:264-271 root_mean_square, :289-296 invert_too_short_consecutive_false,
// :307-329 framewise_rms_energy_vad_decisions, :337-353 remove_silence / lidbox/data/steps.py:183-200 apply_vad.
#include "common.cuh"
#include <math.h>

namespace lbx {

// rms[row] = sqrt(mean(x[row, :]^2)); one warp per row (a VAD frame or any rank-2 row)
__global__ void __launch_bounds__(256) row_rms_kernel(const float* __restrict__ x, long long rows, long long row_pitch,
                                                     int len, long long rows_per_group, long long group_pitch,
                                                     float* __restrict__ out) {
  LBX_PDL_SYNC();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // row r of group g starts at g*group_pitch + (r % rows_per_group)*row_pitch  (frames of utterance g)
  const long long g = row / rows_per_group, r = row - g * rows_per_group;
  const float* p = x + g * group_pitch + r * row_pitch;
  float s = 0.0f;
  for (int i = lane; i < len; i += 32) {
    const float v = p[i];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = sqrtf(s / (float)len);
}

// per utterance: mean RMS -> threshold -> decisions -> too-short non-speech runs are flipped back to speech
__global__ void __launch_bounds__(256) vad_decide_kernel(const float* __restrict__ rms, long long F, float strength,
                                                        float min_rms_threshold, long long min_len,
                                                        unsigned char* __restrict__ dec) {
  LBX_PDL_SYNC();
  __shared__ float red[256];
  const long long b = blockIdx.x;
  const float* r = rms + b * F;
  unsigned char* d = dec + b * F;
  float s = 0.0f;
  for (long long i = threadIdx.x; i < F; i += 256) s += r[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float mean_rms = red[0] / (float)F;
  const float thr = strength * fmaxf(min_rms_threshold, mean_rms);
  for (long long i = threadIdx.x; i < F; i += 256) d[i] = r[i] > thr ? 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0 && min_len > 0) {
    // run-length pass (audio.py:289-296): a run of False shorter than min_len becomes True
    long long i = 0;
    while (i < F) {
      if (d[i]) { ++i; continue; }
      long long j = i;
      while (j < F && !d[j]) ++j;
      if (j - i < min_len)
        for (long long k = i; k < j; ++k) d[k] = 1;
      i = j;
    }
  }
}

// exclusive scan of the decisions of every utterance -> destination frame index, and the voiced length in samples
__global__ void __launch_bounds__(256) vad_scan_kernel(const unsigned char* __restrict__ dec, long long B, long long F,
                                                      int frame_len, long long* __restrict__ offsets,
                                                      long long* __restrict__ out_len) {
  LBX_PDL_SYNC();
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long n = 0;
  for (long long f = 0; f < F; ++f) {
    offsets[b * F + f] = n;
    n += dec[b * F + f] ? 1 : 0;
  }
  out_len[b] = n * frame_len;
}

__global__ void __launch_bounds__(128) vad_copy_kernel(const float* __restrict__ sig, long long N, long long F,
                                                      int frame_len, const unsigned char* __restrict__ dec,
                                                      const long long* __restrict__ offsets, float* __restrict__ out) {
  LBX_PDL_SYNC();
  const long long f = blockIdx.x, b = blockIdx.y;
  if (!dec[b * F + f]) return;
  const float* src = sig + b * N + f * frame_len;
  float* dst = out + b * N + offsets[b * F + f] * frame_len;
  for (int i = threadIdx.x; i < frame_len; i += 128) dst[i] = src[i];
}

}  // namespace lbx

using namespace lbx;

extern "C" {

int lbx_row_rms_f32(const float* x, long long rows, int len, float* out, void* stream) {
  LBX_CHECK_ARG(rows >= 0 && len >= 1, "bad shape");
  if (rows == 0) return LBX_OK;
  LBX_CHECK_ARG(x && out, "NULL pointer argument");
  LBX_LAUNCH_PDL(row_rms_kernel, dim3((unsigned)ceil_div(rows, 8)), dim3(256), 0, (cudaStream_t)stream, x, rows,
                 (long long)len, len, rows, (long long)0, out);
  return LBX_OK;
}

int lbx_rms_vad_f32(const float* sig, long long B, long long N, int frame_step, float strength,
                    float min_rms_threshold, long long min_non_speech_frames, unsigned char* decisions, float* rms_ws,
                    void* stream) {
  LBX_CHECK_ARG(B >= 0 && N >= 0 && frame_step >= 1, "bad shape");
  const long long F = N / frame_step;                         // tf.signal.frame(step == length, pad_end=False)
  if (B == 0 || F == 0) return LBX_OK;
  LBX_CHECK_ARG(sig && decisions && rms_ws, "NULL pointer argument");
  LBX_CHECK_ARG(B <= 2147483647LL && min_non_speech_frames >= 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  LBX_LAUNCH_PDL(row_rms_kernel, dim3((unsigned)ceil_div(B * F, 8)), dim3(256), 0, st, sig, B * F, (long long)frame_step,
                 frame_step, F, N, rms_ws);
  LBX_LAUNCH_PDL(vad_decide_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)rms_ws, F, strength,
                 min_rms_threshold, min_non_speech_frames, decisions);
  return LBX_OK;
}

int lbx_vad_compact_f32(const float* sig, long long B, long long N, int frame_len, const unsigned char* decisions,
                        long long F, float* out, long long* out_len, long long* offsets_ws, void* stream) {
  LBX_CHECK_ARG(B >= 0 && N >= 0 && frame_len >= 1 && F >= 0 && F * frame_len <= N, "bad shape");
  if (B == 0) return LBX_OK;
  LBX_CHECK_ARG(out_len != nullptr, "NULL out_len");
  cudaStream_t st = (cudaStream_t)stream;
  if (F == 0) {
    LBX_CUDA(cudaMemsetAsync(out_len, 0, (size_t)B * sizeof(long long), st));
    return LBX_OK;
  }
  LBX_CHECK_ARG(sig && decisions && out && offsets_ws, "NULL pointer argument");
  LBX_CHECK_ARG(B <= 65535 && F <= 2147483647LL, "shape exceeds the launch limits");
  LBX_LAUNCH_PDL(vad_scan_kernel, dim3((unsigned)ceil_div(B, 256)), dim3(256), 0, st, decisions, B, F, frame_len,
                 offsets_ws, out_len);
  LBX_LAUNCH_PDL(vad_copy_kernel, dim3((unsigned)F, (unsigned)B), dim3(128), 0, st, sig, N, F, frame_len, decisions,
                 (const long long*)offsets_ws, out);
  return LBX_OK;
}

}  // extern "C"
