// Shared helpers for libvbt_b200.so (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vbt_b200.h"

namespace vbt {

void set_error(const char* fmt, ...);
void count_launches(long long n);
int ensure_device();  // VBT_OK when an sm_100 device is current, else VBT_ECUDA

#define VBT_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      vbt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                       \
      return VBT_ECUDA;                                                               \
    }                                                                                 \
  } while (0)

#define VBT_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      vbt::set_error(__VA_ARGS__);  \
      return VBT_EINVAL;            \
    }                               \
  } while (0)

// call after every kernel launch: records it and surfaces launch-configuration errors
#define VBT_LAUNCHED(n)                       \
  do {                                        \
    vbt::count_launches(n);                   \
    VBT_CHECK_CUDA(cudaPeekAtLastError());    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch: consecutive kernels of the layer program are launched with
// the stream-serialisation attribute, so the next kernel's CTAs are scheduled while the
// previous kernel drains and run their prologue (weight loads, TMEM allocation, barrier
// init).  Every such kernel calls pdl_wait() before it touches anything a predecessor wrote
// (or that a predecessor may still be reading), then pdl_launch_dependents().
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// high_priority: the launch (and the graph node captured from it) gets the device's greatest
// priority, so its CTAs are dispatched ahead of pending CTAs of ordinary launches from OTHER
// streams -- used for the small late-network kernels, which otherwise queue behind the other
// detection lane's GPU-filling backbone kernels.
inline int greatest_priority() {
  static const int p = [] { int least = 0, greatest = 0; cudaDeviceGetStreamPriorityRange(&least, &greatest); return greatest; }();
  return p;
}

template <typename Arg>
cudaError_t launch_pdl(void (*kernel)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg& arg,
                       bool high_priority = false, bool pdl = true) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (high_priority) {
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = greatest_priority();
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}
#endif

}  // namespace vbt
