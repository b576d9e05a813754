// Depthwise KxK convolution (K = 3 / 5, stride 1 / 2, TF-SAME) on the tcgen05 tensor pipe.
//
// replaces: the DEPTHWISE_CONV_2D ops of the EfficientDet-Lite graph inside tflite_runtime's
// signature_fn(images=...) (odt.py:58-61) -- 80 of Lite0's 208 ops.
//
// Why tensor cores for an op with no channel reduction: on the SIMT pipes a depthwise tap
// costs one dp4a per output channel word and the kernels end up instruction-bound at ~5 %
// of the HBM roofline.  Written as implicit GEMM with BLOCK-DIAGONAL weights, a tap over
// 32 channels is one tcgen05.mma (M = 128 pixels, N = 32, K = 32, kind::i8): 31/32 of the
// multiplies hit zeros, but the tensor pipe has ~100x the int8 rate of dp4a, so the MACs
// disappear from the critical path and the op becomes what it should be -- a stream of
// activations through shared memory with a requantising epilogue.
//
// One CTA = one frame x one pair of 16-channel groups x one band of output rows:
//   1. cp.async the band's input window (zero-point padded) into shared memory as one
//      "plane" per channel group and stride phase: plane[(ey % S, ex % S)][ey / S][ex / S]
//      holds the 16 channel bytes of extended input pixel (ey, ex) = (iy + pad, ix + pad).
//      Within a plane consecutive x are 16 bytes apart, which is exactly the K-major
//      no-swizzle core-matrix layout (8 rows x 16 bytes), so for output position
//      q = y * PW + x (PW = plane width) tap (ky, kx) reads plane position
//      q + (ky / S) * PW + kx / S: a pure shift of the A descriptor's start address.
//   2. one thread issues K*K MMAs per 128-position tile (A = shifted plane view, LBO = the
//      distance between the two groups' planes; B = the tap's diagonal block) accumulating
//      in TMEM, one tcgen05.commit per tile;
//   3. the four warps read the int32 accumulators with tcgen05.ld, requantise and store
//      32 bytes per valid position (positions in the K-1 pad columns are discarded).
#include "model.cuh"
#include "requant.cuh"

namespace {

using vbt::OpRecord;

constexpr int kMaxTiles = 8;          // 128-position tiles per CTA (TMEM: 32 columns each)

struct DwUArgs {
  const int8_t* in; int8_t* out;
  const int8_t* wdiag;                // [pairs][K*K][1024] block-diagonal tap matrices
  const int32_t* bias; const float* mult;
  int B, H, W, Ho, Wo, c_p, pad_top, pad_left, zp_in;
  int K, S, groups, pairs;
  int TH, n_bands, PW, n_mt, rows_alloc, plane_pos, tmem_cols;
  uint32_t inv_pw;                    // ceil(2^32 / PW)
  vbt::Requant rq;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1LL << 24)) __trap();
  }
}

constexpr int kDwThreads = 256;    // 8 warps: warp w owns TMEM lanes 32 * (w % 4) .., channel group w / 4 of the pair

template <bool FAST>
__global__ void __launch_bounds__(kDwThreads) dw_umma_kernel(DwUArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar[kMaxTiles];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) int32_t sBias[32];
  __shared__ __align__(16) float sMult[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int band = blockIdx.x, gp = blockIdx.y, b = blockIdx.z;
  const int ss = a.S - 1;                              // stride 1 -> shift 0, stride 2 -> shift 1
  const int n_planes = a.S * a.S;
  const uint32_t plane_bytes = (uint32_t)a.plane_pos * 16;
  unsigned char* planes = smem;                                        // [plane][group 0/1][pos][16]
  unsigned char* wsm = smem + (size_t)n_planes * 2 * plane_bytes;      // [K*K][1024]
  const int oy0 = band * a.TH;
  const int th = min(a.TH, a.Ho - oy0);
  const int taps = a.K * a.K;

  // ---- prologue on model constants only (overlaps the previous kernel's tail) -------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid < a.n_mt) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[tid])));
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;\n");
  if (tid < 32) {
    const int c = gp * 32 + tid;
    sBias[tid] = c < a.c_p ? a.bias[c] : 0;
    sMult[tid] = c < a.c_p ? a.mult[c] : 0.f;
  }
  {
    const int4* src = reinterpret_cast<const int4*>(a.wdiag + (size_t)gp * taps * 1024);
    for (int i = tid; i < taps * 64; i += kDwThreads) {
      const uint32_t dst = smem_u32(wsm) + (uint32_t)i * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + i));
    }
  }
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();

  // ---- input window -> planes ---------------------------------------------------------------
  {
    const uint32_t zpw = (uint32_t)(a.zp_in & 0xff) * 0x01010101u;
    const int g0 = gp * 2;
    const bool has_g1 = g0 + 1 < a.groups;
    const int8_t* fin = a.in + (size_t)b * a.H * a.W * a.c_p + g0 * 16;
    const int n_slots = a.rows_alloc * a.PW;
    for (int pl = 0; pl < n_planes; ++pl) {
      const int pr = pl >> ss, pc = pl & ss;           // S = 2: pl = pr * 2 + pc; S = 1: 0, 0
      unsigned char* p0 = planes + (size_t)(pl * 2) * plane_bytes;
      for (int i = tid; i < n_slots; i += kDwThreads) {
        const int ly = (int)__umulhi((uint32_t)i, a.inv_pw);
        const int lx = i - ly * a.PW;
        const int iy = (((oy0 + ly) << ss) + pr) - a.pad_top;
        const int ix = ((lx << ss) + pc) - a.pad_left;
        const uint32_t d0 = smem_u32(p0) + (uint32_t)i * 16;
        const uint32_t d1 = d0 + plane_bytes;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
          const int8_t* src = fin + ((size_t)iy * a.W + ix) * a.c_p;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d0), "l"(src));
          if (has_g1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d1), "l"(src + 16));
          else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};\n" ::"r"(d1), "r"(0u));
        } else {                                       // TF-SAME padding carries the zero point
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};\n" ::"r"(d0), "r"(zpw));
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};\n" ::"r"(d1), "r"(has_g1 ? zpw : 0u));
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;\n");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;

  // ---- MMAs: K*K shifted views per 128-position tile ----------------------------------------
  if (tid == 0) {
    // D = S32, A = B = signed int8, K-major, M = 128, N = 32
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t pbase = smem_u32(planes), wbase = smem_u32(wsm);
    for (int mt = 0; mt < a.n_mt; ++mt) {
      int t = 0;
      for (int ky = 0; ky < a.K; ++ky) {
        for (int kx = 0; kx < a.K; ++kx, ++t) {
          const int pl = ((ky & ss) << ss) | (kx & ss);
          const uint32_t shift = (uint32_t)((ky >> ss) * a.PW + (kx >> ss));
          const uint32_t aaddr = pbase + (uint32_t)(pl * 2) * plane_bytes + ((uint32_t)mt * 128 + shift) * 16;
          umma_i8(tmem + (uint32_t)mt * 32, umma_desc(aaddr, plane_bytes, 128),
                  umma_desc(wbase + (uint32_t)t * 1024, 512, 128), idesc, t > 0 ? 1u : 0u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
          smem_u32(&mbar[mt])) : "memory");
    }
  }

  // ---- epilogue: thread t owns position mt*128 + (t % 128) and channel group t / 128 of the pair --
  const int row = tid & 127, half = tid >> 7;
  const bool has_group = gp * 2 + half < a.groups;
  for (int mt = 0; mt < a.n_mt; ++mt) {
    if (warp == 0) mbar_wait(smem_u32(&mbar[mt]), 0);   // one warp polls, the others sleep on the barrier
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)mt * 32 + (uint32_t)half * 16;
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
        "%13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    const int q = mt * 128 + row;
    const int ly = (int)__umulhi((uint32_t)q, a.inv_pw);
    const int lx = q - ly * a.PW;
    if (has_group && ly < th && lx < a.Wo) {
      uint32_t packed[4];
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const int4 bq = *reinterpret_cast<const int4*>(sBias + half * 16 + w4 * 4);
        const float4 mq = *reinterpret_cast<const float4*>(sMult + half * 16 + w4 * 4);
        packed[w4] = a.rq.pack4t<FAST>((int)v[w4 * 4 + 0] + bq.x, (int)v[w4 * 4 + 1] + bq.y, (int)v[w4 * 4 + 2] + bq.z,
                                       (int)v[w4 * 4 + 3] + bq.w, mq.x, mq.y, mq.z, mq.w);
      }
      int8_t* o = a.out + (((size_t)b * a.Ho + oy0 + ly) * a.Wo + lx) * a.c_p + gp * 32 + half * 16;
      *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem),
                 "r"((uint32_t)a.tmem_cols));
  }
}

}  // namespace

namespace vbt {

// Depthwise op on the tensor pipe; *taken = false leaves it to the SIMT kernel in net.cu.
int launch_dw_umma(const vbt_model* m, const OpRecord& op, const int8_t* in, int8_t* out, int B,
                   cudaStream_t st, bool* taken) {
  static const bool enabled = [] { const char* e = getenv("VBT_DW_IMPL"); return !(e && e[0] == 's'); }();
  *taken = false;
  if (!enabled || op.lut_off < 0 || (op.k != 3 && op.k != 5) || (op.stride != 1 && op.stride != 2)) return VBT_OK;
  // Measured on B200 (profiles/, per-op timings): the tensor-pipe form wins 1.5-2x for 5x5 taps
  // (25 MMAs replace 100 dp4a per output word); for 3x3 the channel-word-stationary SIMT kernel
  // in net.cu is as fast or faster, so it keeps those unless VBT_DW_IMPL=umma forces this one.
  static const bool force = [] { const char* e = getenv("VBT_DW_IMPL"); return e && e[0] == 'u'; }();
  if (op.k == 3 && !force) return VBT_OK;
  DwUArgs a;
  a.in = in; a.out = out;
  a.wdiag = reinterpret_cast<const int8_t*>(m->dev_data + op.lut_off);
  a.bias = reinterpret_cast<const int32_t*>(m->dev_data + op.bias_off);
  a.mult = reinterpret_cast<const float*>(m->dev_data + op.scale_off);
  a.B = B; a.H = op.h_in; a.W = op.w_in; a.Ho = op.h_out; a.Wo = op.w_out; a.c_p = op.cout_p;
  a.pad_top = op.pad_top; a.pad_left = op.pad_left; a.zp_in = op.zp_in[0];
  a.K = op.k; a.S = op.stride;
  a.groups = op.cout_p / 16; a.pairs = (a.groups + 1) / 2;
  const int halo = (op.k - 1) >> (op.stride - 1);          // extra plane rows / columns
  a.PW = op.w_out + halo;
  // positions per CTA: small CTAs (2-4 tiles) keep 4-8 of them resident per SM, which hides the
  // serial load -> MMA -> epilogue chain of each far better than two 8-tile CTAs do
  static const int pos_env = [] { const char* e = getenv("VBT_DW_POS"); return e ? atoi(e) : 256; }();
  const int max_pos = std::max(128, std::min(1024, pos_env));
  // ... but a band re-reads K-S input rows, so never fewer than 4 output rows per band
  a.TH = std::max(std::min(op.h_out, 4), std::min(op.h_out, max_pos / a.PW));
  while (a.TH > 1 && (a.TH * a.PW + 127) / 128 > kMaxTiles) --a.TH;
  a.n_bands = (op.h_out + a.TH - 1) / a.TH;
  a.n_mt = (a.TH * a.PW + 127) / 128;
  if (a.n_mt > kMaxTiles) return VBT_OK;                   // a single row wider than the tile budget
  a.rows_alloc = a.TH + halo;
  a.plane_pos = a.n_mt * 128 + halo * a.PW + halo + 8;
  if (a.plane_pos < a.rows_alloc * a.PW) a.plane_pos = a.rows_alloc * a.PW;
  int cols = 32;
  while (cols < a.n_mt * 32) cols <<= 1;
  a.tmem_cols = cols;
  a.inv_pw = (uint32_t)((0x100000000ULL + a.PW - 1) / a.PW);
  a.rq = Requant(op.zp_out, op.act_lo, op.act_hi, op.requant_fast);
  size_t smem = (size_t)op.stride * op.stride * 2 * a.plane_pos * 16 + (size_t)op.k * op.k * 1024;
  if (smem > 200 * 1024) return VBT_OK;
  const size_t cap_ctas = 512 / cols;                      // never more CTAs than TMEM can serve
  smem = std::max(smem, (size_t)228 * 1024 / (cap_ctas + 1));
  static bool attr_set = false;
  if (!attr_set) {
    VBT_CHECK_CUDA(cudaFuncSetAttribute(dw_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VBT_CHECK_CUDA(cudaFuncSetAttribute(dw_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)a.n_bands, (unsigned)a.pairs, (unsigned)B);
  if (a.rq.fast) VBT_CHECK_CUDA(launch_pdl(dw_umma_kernel<true>, grid, dim3(kDwThreads), smem, st, a));
  else VBT_CHECK_CUDA(launch_pdl(dw_umma_kernel<false>, grid, dim3(kDwThreads), smem, st, a));
  *taken = true;
  return VBT_OK;
}

}  // namespace vbt
