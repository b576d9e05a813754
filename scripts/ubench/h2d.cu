// Host -> device ingest patterns for the bilinear resize 1080p -> 320 (only 16 of every 27 rows and 2 of every 6
// pixels of a touched row are read): (a) one contiguous copy of the whole batch, (b) 16 strided 2-D copies of whole
// touched rows (what ingest.RowSparseIngest does), (c) 2-D copies that move only the 6 needed bytes of every 18 (width 6, pitch 18).
// build: nvcc -O3 -o h2d h2d.cu ; run: ./h2d
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main() {
  const int B = 64, H = 1080, W = 1920, RB = W * 3;          // row bytes 5760
  const size_t frame = (size_t)H * RB;
  unsigned char *h, *d;
  CK(cudaMallocHost(&h, B * frame));
  CK(cudaMalloc(&d, B * frame));
  for (size_t i = 0; i < B * frame; i += 4096) h[i] = (unsigned char)i;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // the 16 touched rows of every period of 27 source rows (oy = 0..7: rows floor((oy + .5) * 3.375 - .5), + 1)
  int rows[16], n = 0;
  for (int oy = 0; oy < 8; ++oy) { int y = (int)((oy + 0.5) * 3.375 - 0.5); rows[n++] = y; rows[n++] = y + 1; }
  const int periods = H / 27;                                 // 40
  auto timeit = [&](const char* name, auto fn, double bytes) {
    for (int w = 0; w < 2; ++w) fn();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) fn();
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("%-44s %8.3f ms per 64 frames  %7.1f GB/s moved  -> %8.0f frames/s\n", name, ms, bytes / ms / 1e6, 64e3 / ms);
  };
  timeit("(a) whole frames, one copy", [&] { CK(cudaMemcpyAsync(d, h, B * frame, cudaMemcpyHostToDevice, st)); }, (double)B * frame);
  timeit("(b) 16 x 2-D copies of touched rows", [&] {
    for (int k = 0; k < 16; ++k)      // rows k of every period, all frames: pitch 27 rows, height = B * 40 periods
      CK(cudaMemcpy2DAsync(d + (size_t)k * RB, (size_t)16 * RB, h + (size_t)rows[k] * RB, (size_t)27 * RB, RB, (size_t)B * periods,
                           cudaMemcpyHostToDevice, st));
  }, (double)16 * RB * B * periods);
  // (c) as 2-D copies: width 6, pitch 18, height 320 -> one source row per copy; time 2048 of them and scale
  {
    const int sample = 2048;
    for (int w = 0; w < 64; ++w)
      CK(cudaMemcpy2DAsync(d, 6, h + 6, 18, 6, 320, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < sample; ++i)
      CK(cudaMemcpy2DAsync(d + (size_t)i * 1920, 6, h + (size_t)(i % (B * H)) * RB + 6, 18, 6, 320, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("(c') 2-D copy of one row's 320 x 6 bytes: %.2f us per row -> %.0f frames/s if every touched row were copied so\n", 1e3 * ms / sample,
           1e3 / (ms / sample * 640));
  }
  // (d) wider segments: 2-D copy width 6 * 320 pitch 18 * 320?  no -- (d) tall skinny 2-D copy over MANY rows at once:
  // width 6, pitch 18, height 320 * rows: consecutive source rows are contiguous (5760 = 320 * 18), so one 2-D copy
  // covers any run of consecutive rows; touched rows come in pairs -> runs of 2 rows: 320 copies per frame.  Time one
  // frame-sized run to see the engine's rate on 6-byte segments:
  {
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    CK(cudaMemcpy2DAsync(d, 6, h + 6, 18, 6, (size_t)320 * H * 8, cudaMemcpyHostToDevice, st));     // 8 whole frames
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("(d) one 2-D copy, width 6 pitch 18, 8 frames of segments: %.3f ms -> %.2f GB/s of useful bytes, %.0f frames/s at 640 rows per frame\n",
           ms, 6.0 * 320 * H * 8 / ms / 1e6, 8e3 / ms * 1080.0 / 640.0);
  }
  return 0;
}
