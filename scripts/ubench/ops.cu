// Issue-throughput micro-benchmark of the integer / conversion instructions the requantising epilogues use
// (sm_100a).  One CTA per SM, 512 threads, 16 independent chains per thread; prints thread-ops per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ops ops.cu ; run: ./ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) bench(uint32_t* out, long long* cyc, int iters, uint32_t a0, uint32_t b0, uint32_t c0) {
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = a0 + threadIdx.x * 16 + j;
  __shared__ unsigned char sm[16384];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) v[j] = (uint32_t)__viaddmin_s32_relu((int)v[j], (int)b0, (int)c0);
      if (OP == 1) v[j] = __viaddmin_s16x2_relu(v[j], b0, c0);
      if (OP == 2) v[j] = (uint32_t)max(min((int)v[j] + (int)b0, (int)c0), 0);
      if (OP == 3) v[j] = __vminu2(__vmaxu2(v[j], b0), c0);
      if (OP == 4) v[j] = __float_as_uint(__int2float_rn((int)v[j]));
      if (OP == 5) v[j] = __float_as_uint(__fmul_rn(__uint_as_float(v[j]), __uint_as_float(b0)));
      if (OP == 6) v[j] = __byte_perm(v[j], b0, 0x5410 + (c0 & 1));
      if (OP == 7) v[j] = (uint32_t)__dp4a((int)v[j], (int)b0, (int)v[j]);
      if (OP == 8) v[j] = v[j] + b0;
      if (OP == 9) { sm[(threadIdx.x + j * 36 * 32 + (v[j] & 0)) & 16383] = (unsigned char)v[j]; }
      if (OP == 10) { reinterpret_cast<uint32_t*>(sm)[(threadIdx.x + j * 512 + (v[j] & 0)) & 4095] = v[j]; }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[threadIdx.x];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc, int threads) {
  const int iters = 2000;
  bench<OP><<<148, threads>>>(out, cyc, iters, 1u, 0x10003u, 0x7fff7fffu);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-28s %4d threads: %7.1f thread-ops/clk/SM\n", name, threads, (double)iters * 16 * threads / avg);
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int threads : {256, 512}) {
    run<0>("VIADDMNMX.RELU (s32)", out, cyc, threads);
    run<1>("VIADDMNMX.S16x2.RELU", out, cyc, threads);
    run<2>("IADD + IMNMX + IMNMX", out, cyc, threads);
    run<3>("VIMNMX.U16x2 pair", out, cyc, threads);
    run<4>("I2FP.F32.S32", out, cyc, threads);
    run<5>("FMUL", out, cyc, threads);
    run<6>("PRMT", out, cyc, threads);
    run<7>("IDP.4A", out, cyc, threads);
    run<8>("IADD", out, cyc, threads);
    run<9>("STS.U8 (lanes = bytes)", out, cyc, threads);
    run<10>("STS.32", out, cyc, threads);
  }
  return 0;
}
