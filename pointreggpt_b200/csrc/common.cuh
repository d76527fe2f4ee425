// Shared host/device helpers for libprg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>

#include "../../include/prg.h"

namespace prg {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define PRG_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      prg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                        \
      return PRG_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define PRG_CHECK_ARG(cond, msg)              \
  do {                                        \
    if (!(cond)) {                            \
      prg::set_error("bad argument: %s", msg); \
      return PRG_ERR_ARG;                     \
    }                                         \
  } while (0)

// Every kernel launch in the library goes through this so prg_launch_count() is exact.
#define PRG_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    prg::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      prg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                               \
      return PRG_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Selects the device that owns `ptr` for the lifetime of the guard (entry points take raw device
// pointers and a stream of that device; the caller's current device may be another GPU).
struct PtrDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit PtrDeviceGuard(const void* ptr) {
    cudaPointerAttributes a;
    if (ptr == nullptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    if (prev != a.device && cudaSetDevice(a.device) == cudaSuccess) switched = true;
  }
  ~PtrDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// ---- programmatic dependent launch ---------------------------------------------------------
// Every kernel of a network evaluation starts with pdl_trigger() (the next kernel of the stream may
// be scheduled as soon as this grid's CTAs are all running or gone) and calls pdl_wait() before it
// touches anything a predecessor wrote (until then: barrier initialisation, TMEM allocation,
// tensor-map prefetch, weight / bias loads).  So the prologue of kernel i + 1 overlaps the tail of
// kernel i -- 127 times per sampler step, inside the step's CUDA graph as programmatic edges.
// A kernel launched without the attribute sees both instructions as no-ops.
extern bool g_pdl_enabled;     // PRG_NO_PDL=1 or a profiled evaluation: plain stream order
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace prg
