// Error reporting, launch accounting and device check shared by every entry point.
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace vbt {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (libvbt_b200 has no CPU path)", cudaGetErrorString(e));
    return VBT_ECUDA;
  }
  static thread_local int checked_dev = -1;
  if (checked_dev == dev) return VBT_OK;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    set_error("device %d is not sm_100 (compute capability major %d): libvbt_b200 is built "
              "for sm_100a only", dev, major);
    return VBT_ECUDA;
  }
  checked_dev = dev;
  return VBT_OK;
}

}  // namespace vbt

extern "C" {

int vbt_abi_version(void) { return 1; }

const char* vbt_last_error(void) { return vbt::g_error; }

long long vbt_launch_count(void) { return vbt::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
