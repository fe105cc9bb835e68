// Shared helpers for the lidbox_b200 C-ABI translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "lidbox_b200.h"

namespace lbx {

// thread-local error message returned by lbx_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define LBX_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return lbx::set_error(LBX_EINVAL, __VA_ARGS__);  \
  } while (0)

#define LBX_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return lbx::set_error(LBX_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
  } while (0)

#define LBX_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return lbx::set_error(LBX_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                \
    lbx::count_launch();                                                                        \
  } while (0)

// Programmatic dependent launch: every kernel of the library starts with LBX_PDL_SYNC() (wait for the previous
// kernel's memory, then let the next kernel start its own launch/prologue) and is launched through launch_pdl().
extern int g_use_pdl;
#define LBX_PDL_SYNC()                                             \
  do {                                                             \
    asm volatile("griddepcontrol.wait;" ::: "memory");             \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define LBX_LAUNCH_PDL(...)                                                                          \
  do {                                                                                               \
    cudaError_t _le = lbx::launch_pdl(__VA_ARGS__);                                                  \
    if (_le != cudaSuccess)                                                                          \
      return lbx::set_error(LBX_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_le),  \
                            __FILE__, __LINE__);                                                     \
    lbx::count_launch();                                                                             \
  } while (0)

static inline bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
static inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace lbx
