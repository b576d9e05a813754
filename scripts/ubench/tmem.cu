// TMEM read / write throughput per SM (sm_100a): one CTA per SM, W warps, each tcgen05.ld / tcgen05.st 32x32b.x16
// (2 KB per warp instruction) in a loop.  Prints bytes per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld + wait each; 1: two lds in flight; 2: st + wait each; 3: x32 ld
__global__ void bench(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t t = base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i & 1) * 32);
    if (MODE == 0 || MODE == 1) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(t + col));
      if (MODE == 1)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(t + col + 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      acc += v[0] + v[15] + (MODE == 1 ? v[16] + v[31] : 0);
    } else if (MODE == 2) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(t + col),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base_s), "r"(512u));
}

template <int MODE>
void run(const char* name, uint32_t* out, long long* cyc, int threads) {
  const int iters = 4000;
  bench<MODE><<<148, threads>>>(out, cyc, iters);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed\n", name); return; }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double bytes = (double)iters * (threads / 32) * 2048 * (MODE == 1 ? 2 : 1);
  printf("%-30s %2d warps: %7.1f B/clk/SM, %6.1f clk per iteration\n", name, threads / 32, bytes / avg, avg / iters);
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int threads : {32, 128, 256}) {
    run<0>("tcgen05.ld x16 + wait", out, cyc, threads);
    run<1>("2 x tcgen05.ld x16 + wait", out, cyc, threads);
    run<2>("tcgen05.st x16 + wait", out, cyc, threads);
  }
  return 0;
}
