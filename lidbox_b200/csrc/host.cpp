// Host-side pieces of the C-ABI: error reporting, integer frame arithmetic, the (bug-compatible) mel table.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>

namespace lbx {

std::atomic<long long> g_launch_count{0};
int g_use_pdl = 1;

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace lbx

extern "C" {

const char* lbx_last_error(void) { return lbx::error_buffer(); }
int lbx_version(void) { return 100; }
long long lbx_launch_count(void) { return lbx::g_launch_count.load(); }

// programmatic dependent launch (on by default; LBX_PDL=0 in the environment of the Python host disables it)
int lbx_set_pdl(int enabled) {
  lbx::g_use_pdl = enabled ? 1 : 0;
  return LBX_OK;
}

// lidbox/features/audio.py:185-189: tf.cast(tf.cast(sr, f32) * 1e-3 * tf.cast(ms, f32), i32); left-to-right fp32
int lbx_ms_to_frames(int sample_rate, int ms) {
  volatile float a = (float)sample_rate * 1e-3f;
  volatile float b = a * (float)ms;
  return (int)b;
}

long long lbx_num_frames(long long n_samples, int frame_length, int frame_step) {
  if (frame_length < 1 || frame_step < 1 || n_samples < frame_length) return 0;
  return 1 + (n_samples - frame_length) / frame_step;
}

// lidbox/features/mel_ops.py:28-75 with its _linspace (:11-16) that divides by num instead of num-1.
int lbx_mel_weight_matrix(int n_mel, int n_bins, int sample_rate, float lower_edge_hertz, float upper_edge_hertz,
                          float* W_host) {
  LBX_CHECK_ARG(n_mel >= 1 && n_bins >= 2 && W_host != nullptr, "bad mel table arguments");
  const float nyquist = (float)sample_rate / 2.0f;
  auto hz_to_mel = [](float f) -> float {
    volatile float r = f / 700.0f;
    volatile float a = 1.0f + r;
    volatile float l = logf(a);
    return 1127.0f * l;
  };
  const float mel_lo = hz_to_mel(lower_edge_hertz), mel_hi = hz_to_mel(upper_edge_hertz);
  std::vector<float> edges(n_mel + 2);
  for (int j = 0; j < n_mel + 2; ++j) {
    volatile float d = mel_hi - mel_lo;
    volatile float num = d * (float)j;
    volatile float q = num / (float)(n_mel + 2);
    edges[j] = mel_lo + q;
  }
  for (int m = 0; m < n_mel; ++m) W_host[m] = 0.0f;   // DC row re-added by tf.pad (mel_ops.py:74-75)
  for (int k = 1; k < n_bins; ++k) {
    volatile float num = nyquist * (float)k;           // 0 + (nyq - 0) * k / n_bins
    volatile float hz = num / (float)n_bins;
    const float mel = hz_to_mel(hz);
    for (int m = 0; m < n_mel; ++m) {
      const float lower = edges[m], center = edges[m + 1], upper = edges[m + 2];
      volatile float ls = (mel - lower) / (center - lower);
      volatile float us = (upper - mel) / (upper - center);
      const float w = fminf(ls, us);
      W_host[(size_t)k * n_mel + m] = (w > 0.0f) ? w : 0.0f;   // degenerate (NaN) slopes are zeroed
    }
  }
  return LBX_OK;
}

int lbx_mel_pack_bands(const float* W_host, int n_bins, int n_mel, int* start_host, int* len_host, int* off_host,
                       float* packed_host) {
  LBX_CHECK_ARG(W_host && start_host && len_host && off_host && packed_host && n_bins >= 1 && n_mel >= 1,
                "bad pack arguments");
  int n = 0;
  for (int m = 0; m < n_mel; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < n_bins; ++k) {
      if (W_host[(size_t)k * n_mel + m] != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    off_host[m] = n;
    if (first < 0) {
      start_host[m] = 0;
      len_host[m] = 0;
      continue;
    }
    start_host[m] = first;
    len_host[m] = last - first + 1;
    for (int k = first; k <= last; ++k) packed_host[n++] = W_host[(size_t)k * n_mel + m];
  }
  return n;
}

}  // extern "C"
