// K1 -- fused frame preprocessing: BGR->RGB swap + bilinear resize to SxS + truncating
// cast, uint8 HWC in, uint8 HWC out.
//
// replaces: cv2.cvtColor(frame, COLOR_BGR2RGB) (track.py:171) and odt.preprocess_image
// (odt.py:10-19): tf.image.resize(img, (S,S)) -- bilinear, half-pixel centres, no
// antialias, aspect ratio NOT preserved -- followed by tf.cast(..., uint8) (truncation).
// The model's (x-127)/128 normalisation is folded into the stem conv's input zero point,
// exactly as the exported int8 graph does.
//
// HBM-bound.  Only 2 source rows feed one output row, so the kernel is organised by
// output row: a CTA stages the two source rows it needs in shared memory with 1-D bulk
// async copies (TMA, cp.async.bulk + mbarrier; rows are 16-byte multiples for 1080p) and
// double-buffers the next row pair behind the arithmetic; output rows are written with
// coalesced 32-bit stores.  Algorithmic bytes per frame: H*W*3 read + S*S*3 written
// (the touched-row lower bound is 2*S*W*3).
// fp32 in the reference's operation order, -fmad=false so the truncation boundary matches.
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

struct Interp { int lo, hi; float lerp; };

__device__ __forceinline__ Interp interp_of(int o, float scale, int in_size) {
  // tf compute_interpolation_weights with the half-pixel scaler
  const float in = ((float)o + 0.5f) * scale - 0.5f;
  const float in_f = floorf(in);
  Interp r;
  r.lo = max((int)in_f, 0);
  r.hi = min((int)ceilf(in), in_size - 1);
  r.lerp = in - in_f;
  return r;
}

__device__ __forceinline__ uint8_t lerp_px(float tl, float tr, float bl, float br, float xl,
                                           float yl) {
  const float top = tl + (tr - tl) * xl;
  const float bottom = bl + (br - bl) * xl;
  return (uint8_t)(top + (bottom - top) * yl);      // tf.cast float32 -> uint8 truncates
}

constexpr int kThreads = 256;
constexpr int kStages = 2;

// Persistent CTAs walk output rows (b, oy).  Row bytes must be a multiple of 16 and the
// frame base 16-byte aligned (true for W*3 % 16 == 0, e.g. 1920x1080): TMA path.
// row_map != nullptr: `frames` is a compacted row table ([B][rows_pf][W*3], only the rows the
// resize touches, see vbt_copy_rows_h2d) and row_map[y] is the table row of source row y.
__global__ void __launch_bounds__(kThreads) preprocess_tma_kernel(
    const uint8_t* __restrict__ frames, int B, int H, int W, int swap_rb,
    uint8_t* __restrict__ out, int S, const int32_t* __restrict__ row_map, int rows_pf) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kStages];
  const int row_bytes = W * 3;
  const int stage_bytes = 2 * row_bytes;
  const float sy = (float)H / (float)S, sx = (float)W / (float)S;
  const int total_rows = B * S;
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int row, int stage) {
    const int b = row / S, oy = row % S;
    const Interp iy = interp_of(oy, sy, H);
    const uint8_t* base = frames + (size_t)b * (row_map ? rows_pf : H) * row_bytes;
    const int r_lo = row_map ? __ldg(row_map + iy.lo) : iy.lo;
    const int r_hi = row_map ? __ldg(row_map + iy.hi) : iy.hi;
    unsigned char* dst = smem + (size_t)stage * stage_bytes;
    mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
    bulk_g2s(dst, base + (size_t)r_lo * row_bytes, (uint32_t)row_bytes, &full[stage]);
    bulk_g2s(dst + row_bytes, base + (size_t)r_hi * row_bytes, (uint32_t)row_bytes, &full[stage]);
  };

  int it = 0;
  const int first = blockIdx.x;
  if (tid == 0 && first < total_rows) issue(first, 0);
  for (int row = first; row < total_rows; row += gridDim.x, ++it) {
    const int stage = it % kStages;
    const int next = row + gridDim.x;
    if (tid == 0 && next < total_rows) issue(next, (it + 1) % kStages);
    mbar_wait(&full[stage], (uint32_t)((it / kStages) & 1));
    const unsigned char* top = smem + (size_t)stage * stage_bytes;
    const unsigned char* bot = top + row_bytes;
    const int b = row / S, oy = row % S;
    const float yl = interp_of(oy, sy, H).lerp;
    uint8_t* orow = out + ((size_t)b * S + oy) * S * 3;
    // 4 output bytes per thread per step: S*3 is a multiple of 4 for S in {320,384,448}
    for (int w4 = tid; w4 * 4 < S * 3; w4 += kThreads) {
      uint32_t packed = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = w4 * 4 + k;
        const int ox = e / 3, c = e - ox * 3;
        const int ci = swap_rb ? 2 - c : c;
        const Interp ix = interp_of(ox, sx, W);
        const uint8_t v = lerp_px((float)top[ix.lo * 3 + ci], (float)top[ix.hi * 3 + ci],
                                  (float)bot[ix.lo * 3 + ci], (float)bot[ix.hi * 3 + ci],
                                  ix.lerp, yl);
        packed |= (uint32_t)v << (8 * k);
      }
      *reinterpret_cast<uint32_t*>(orow + (size_t)w4 * 4) = packed;
    }
    __syncthreads();     // everyone is done with this stage before it is refilled
  }
}

// Generic path for frames whose rows are not 16-byte multiples (e.g. the 416x416 test
// images): direct read-only loads, one output byte triple per thread.
__global__ void preprocess_direct_kernel(const uint8_t* __restrict__ frames, int B, int H, int W,
                                         int swap_rb, uint8_t* __restrict__ out, int S) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * S * S;
  if (i >= total) return;
  const int ox = (int)(i % S), oy = (int)((i / S) % S), b = (int)(i / ((size_t)S * S));
  const float sy = (float)H / (float)S, sx = (float)W / (float)S;
  const Interp iy = interp_of(oy, sy, H), ix = interp_of(ox, sx, W);
  const uint8_t* base = frames + (size_t)b * H * W * 3;
  const uint8_t* r0 = base + (size_t)iy.lo * W * 3;
  const uint8_t* r1 = base + (size_t)iy.hi * W * 3;
  uint8_t* o = out + i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int ci = swap_rb ? 2 - c : c;
    o[c] = lerp_px((float)__ldg(r0 + ix.lo * 3 + ci), (float)__ldg(r0 + ix.hi * 3 + ci),
                   (float)__ldg(r1 + ix.lo * 3 + ci), (float)__ldg(r1 + ix.hi * 3 + ci), ix.lerp,
                   iy.lerp);
  }
}

}  // namespace

static int launch_preprocess(const uint8_t* dev_frames, int B, int H, int W, int swap_rb, uint8_t* dev_out,
                             int S, const int32_t* dev_row_map, int rows_pf, cudaStream_t st) {
  const int row_bytes = W * 3;
  const size_t smem = (size_t)kStages * 2 * row_bytes;
  const bool tma_ok = (row_bytes % 16 == 0) && (((size_t)H * row_bytes) % 16 == 0) &&
                      ((uintptr_t)dev_frames % 16 == 0) && ((S * 3) % 4 == 0) &&
                      ((uintptr_t)dev_out % 4 == 0) && smem <= 200 * 1024;
  if (tma_ok) {
    static thread_local size_t smem_set = 0;
    if (smem_set < smem) {
      VBT_CHECK_CUDA(cudaFuncSetAttribute(preprocess_tma_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = (int)((200 * 1024) / (smem + 1024));
    int grid = sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (grid > B * S) grid = B * S;
    preprocess_tma_kernel<<<grid, kThreads, smem, st>>>(dev_frames, B, H, W, swap_rb, dev_out, S,
                                                        dev_row_map, rows_pf);
  } else {
    VBT_REQUIRE(dev_row_map == nullptr, "vbt_preprocess_rows_u8: rows of %d bytes are not 16-byte multiples",
                row_bytes);
    const size_t total = (size_t)B * S * S;
    preprocess_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dev_frames, B, H, W,
                                                                             swap_rb, dev_out, S);
  }
  VBT_LAUNCHED(1);
  return VBT_OK;
}

extern "C" int vbt_preprocess_u8(const uint8_t* dev_frames, int B, int H, int W, int swap_rb,
                                 uint8_t* dev_out, int S, void* stream) {
  VBT_REQUIRE(dev_frames && dev_out, "vbt_preprocess_u8: null pointer");
  VBT_REQUIRE(B > 0 && H > 0 && W > 0 && S > 0, "vbt_preprocess_u8: B=%d H=%d W=%d S=%d", B, H, W, S);
  if (int rc = vbt::ensure_device()) return rc;
  return launch_preprocess(dev_frames, B, H, W, swap_rb, dev_out, S, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int vbt_preprocess_rows_u8(const uint8_t* dev_rows, int B, int H, int W, int rows_per_frame,
                                      const int32_t* dev_row_map, int swap_rb, uint8_t* dev_out, int S,
                                      void* stream) {
  VBT_REQUIRE(dev_rows && dev_row_map && dev_out, "vbt_preprocess_rows_u8: null pointer");
  VBT_REQUIRE(B > 0 && H > 0 && W > 0 && S > 0 && rows_per_frame > 0,
              "vbt_preprocess_rows_u8: B=%d H=%d W=%d S=%d rows=%d", B, H, W, S, rows_per_frame);
  if (int rc = vbt::ensure_device()) return rc;
  return launch_preprocess(dev_rows, B, H, W, swap_rb, dev_out, S, dev_row_map, rows_per_frame,
                           (cudaStream_t)stream);
}

extern "C" int vbt_copy_rows_h2d(const uint8_t* host_frames, int B, int H, int W, int period,
                                 const int32_t* host_rows, int n_rows, uint8_t* dev_rows, void* stream) {
  VBT_REQUIRE(host_frames && host_rows && dev_rows, "vbt_copy_rows_h2d: null pointer");
  VBT_REQUIRE(B > 0 && H > 0 && W > 0 && period > 0 && H % period == 0 && n_rows > 0 && n_rows <= period,
              "vbt_copy_rows_h2d: B=%d H=%d W=%d period=%d rows=%d", B, H, W, period, n_rows);
  const size_t row_bytes = (size_t)W * 3;
  const size_t periods = (size_t)(H / period) * B;     // frames are contiguous: one long 2-D array
  for (int j = 0; j < n_rows; ++j) {
    VBT_REQUIRE(host_rows[j] >= 0 && host_rows[j] < period, "vbt_copy_rows_h2d: row %d outside the period", host_rows[j]);
    VBT_CHECK_CUDA(cudaMemcpy2DAsync(dev_rows + (size_t)j * row_bytes, (size_t)n_rows * row_bytes,
                                     host_frames + (size_t)host_rows[j] * row_bytes, (size_t)period * row_bytes,
                                     row_bytes, periods, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  return VBT_OK;
}
