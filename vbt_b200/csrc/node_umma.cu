// Fused BiFPN node / head stage:  [quantised sum of resampled inputs ->] depthwise 3x3 ->
// pointwise 1x1, one kernel, intermediates resident in shared memory and TMEM.
//
// replaces: the ADD (+ ReLU6) -> DEPTHWISE_CONV_2D -> CONV_2D 1x1 triples of every BiFPN node
// and the DEPTHWISE_CONV_2D -> CONV_2D pairs of the class / box heads inside tflite_runtime's
// signature_fn(images=...) (odt.py:58-61).  In EfficientDet-Lite0 that is 136 of the 208
// ops, all on 3x3 ... 40x40 maps where a stand-alone kernel is pure launch latency.
//
// One CTA = one frame x one band of rows of the level, ALL channels (C <= 128):
//   1. fill: the band's input window (one halo row / column, zero-point padded) is written to
//      shared-memory planes, one per 16-channel group, [position][16 B] with
//      position = y * PW + x.  Either a copy of one tensor (heads) or the BiFPN fusion
//      computed on the fly: nearest-neighbour up-sampling / 3x3 s2 max-pool down-sampling of
//      up to three inputs, integer rescale, sum, ReLU6 clamp -- the arithmetic of add_kernel.
//      A 3-input sum arrives as the exported graphs hold it, ADD(ADD(a, b), c): the inner sum is
//      requantised to its own int8 tensor (in registers) before the outer one uses it;
//   2. depthwise: as dw_umma.cu -- per 128-position tile and channel-group pair, nine
//      tcgen05.mma with block-diagonal weights whose A descriptors are shifted views of the
//      planes; accumulators in TMEM; the epilogue requantises into a second set of planes;
//   3. pointwise: those planes ARE a K-major core-matrix A operand (8 positions x 16 B per
//      core matrix, the next channel group one plane further), so the 1x1 conv is
//      ceil(groups/2) tcgen05.mma per tile against the [Cout][C] weights, reusing the TMEM
//      columns; the epilogue requantises (+ LOGISTIC LUT for the class head) and stores.
// bf16 heads (BASELINE configs[3], OpRecord.pw_dtype = 1): the pointwise stage runs as
// tcgen05.mma.kind::f16 on bf16 operands with fp32 accumulators.  The operands are the SAME
// integers (|x| <= 128, |w| <= 127: exact in bf16; every product and every partial sum of
// <= 128 terms stays below 2^24: exact in fp32), so the accumulators -- and therefore the
// outputs -- equal the int8 path's bit for bit; only the tensor-pipe data type changes.
#include "model.cuh"
#include "requant.cuh"

namespace {

using vbt::OpRecord;

constexpr int kMaxTiles = 4;       // 128-position tiles per CTA
constexpr int kMaxGroups = 8;      // 16-channel groups (C <= 128)

struct NodeArgs {
  // fill
  const int8_t* in[3];
  int n_in, fused_add;             // fused_add = 0: plain copy of in[0]
  int in_h[3], in_w[3], resample[3], zp_in[3], add_mult[3];
  int add_shift, add_zp, add_lo, add_hi;
  int add_init;                    // rounding constant minus sum_i zp_i * mult_i (the zero points folded out)
  // inner ADD of a two-level sum ADD(ADD(in[0], in[1]), in[2]); pre_n = 0: single-level sum
  int pre_n, pre_mult[2], pre_init, pre_shift, pre_zp, pre_lo, pre_hi, pre_out_mult;
  // geometry (depthwise 3x3 stride 1: output size = input size)
  int B, H, W, c_p, groups, pairs, kch_pad;
  // depthwise
  const int8_t* wdiag; const int32_t* dw_bias; const float* dw_mult; vbt::Requant dw_rq; int dw_zp_in;
  // pointwise
  const int8_t* pw_w; const int32_t* pw_bias; const float* pw_mult; vbt::Requant pw_rq; const int8_t* lut;
  int cout, cout_p, pw_bf16;
  int8_t* out; int out_pix_stride; long long out_batch_stride, out_elem_offset; int vector_out;
  // tiling
  int TH, n_bands, PW, n_mt, rows_alloc, plane_pos, tmem_cols;
  uint32_t inv_pw;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// 8 int8 (two words) -> 8 bf16: float(v) of |v| <= 128 has its whole significand in the upper half
__device__ __forceinline__ uint4 s8x8_to_bf16(uint32_t w0, uint32_t w1) {
  uint32_t f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[i] = __float_as_uint(__int2float_rn((int)(int8_t)(w0 >> (8 * i))));
    f[4 + i] = __float_as_uint(__int2float_rn((int)(int8_t)(w1 >> (8 * i))));
  }
  return make_uint4(__byte_perm(f[0], f[1], 0x7632), __byte_perm(f[2], f[3], 0x7632),
                    __byte_perm(f[4], f[5], 0x7632), __byte_perm(f[6], f[7], 0x7632));
}
// one elected lane of a warp-uniform region issues the MMAs: under a divergent `if (tid == 0)` ptxas wraps
// every UTCIMMA in an ELECT / R2UR / BRA.U.ANY loop (its operands live in uniform registers) -- measured
// at ~100 cycles per instruction in csrc/mbconv_umma.cu
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, lane = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %2;\n\t@px mov.s32 %1, 1;\n\tmov.s32 %0, rx;\n\t}\n"
      : "+r"(lane), "+r"(pred) : "r"(0xffffffffu));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
               : "memory");
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
// one warp polls the mbarrier, the others sleep on the hardware barrier instead of spinning
// through issue slots the epilogues of co-resident CTAs could use
__device__ __forceinline__ void block_wait(uint32_t bar, uint32_t parity, int warp) {
  if (warp == 0) mbar_wait(bar, parity);
  __syncthreads();
}
__device__ __forceinline__ void st_shared16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ int s8(uint32_t word, int i) { return (int)(int8_t)(word >> (8 * i)); }
// byte J of a word, sign-extended: one PRMT (selector bit 3 replicates the byte's sign)
template <int J>
__device__ __forceinline__ int sext_byte(uint32_t word) {
  uint32_t r;      // prmt.b32 directly: the __byte_perm intrinsic only honours three selector bits
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(word), "n"(J | ((J | 8) << 4) | ((J | 8) << 8) | ((J | 8) << 12)));
  return (int)r;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// one input of the BiFPN fusion at level position (oy, ox), 16 channels from c0
__device__ __forceinline__ uint4 fetch_resampled(const int8_t* base, int b, int oy, int ox, int Ho, int Wo,
                                                 int ih, int iw, int mode, int c_p, int c0) {
  const int8_t* fin = base + (size_t)b * ih * iw * c_p + c0;
  if (mode == vbt::RS_NONE)
    return __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)oy * iw + ox) * c_p));
  if (mode == vbt::RS_UP_NEAREST) {
    const int sy = (oy * ih) / Ho, sx = (ox * iw) / Wo;
    return __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)sy * iw + sx) * c_p));
  }
  const int pt = max((Ho - 1) * 2 + 3 - ih, 0) / 2, pl = max((Wo - 1) * 2 + 3 - iw, 0) / 2;
  uint4 m = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - pt + ky;
    if (iy < 0 || iy >= ih) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - pl + kx;
      if (ix < 0 || ix >= iw) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)iy * iw + ix) * c_p));
      m.x = __vmaxs4(m.x, v.x); m.y = __vmaxs4(m.y, v.y);
      m.z = __vmaxs4(m.z, v.z); m.w = __vmaxs4(m.w, v.w);
    }
  }
  return m;
}

constexpr int kThreads = 256;      // 8 warps: warp w owns TMEM lanes 32 * (w % 4) ..., column half w / 4

__global__ void __launch_bounds__(kThreads) node_umma_kernel(NodeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar_dw[kMaxTiles], mbar_pw[kMaxTiles];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) int32_t sDwBias[kMaxGroups * 16];
  __shared__ __align__(16) float sDwMult[kMaxGroups * 16];
  __shared__ __align__(16) int32_t sPwBias[kMaxGroups * 16];
  __shared__ __align__(16) float sPwMult[kMaxGroups * 16];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, half = tid >> 7;            // tile row (= TMEM lane) and column half
  const int band = blockIdx.x, b = blockIdx.y;
  const int G = a.groups, Ge = a.kch_pad;                 // Ge: groups rounded up to even
  const uint32_t in_plane = (uint32_t)a.plane_pos * 16;   // bytes per input plane
  const uint32_t mid_plane = (uint32_t)a.n_mt * 128 * 16; // bytes per depthwise-output plane
  unsigned char* pin = smem;                                        // [Ge][plane_pos][16]
  unsigned char* pmid = pin + (size_t)Ge * in_plane;                // [Ge][n_mt*128][16]
  // bf16 pointwise stage: one plane per 8 channels ([2G][n_mt*128][16 B]) and bf16 weights
  const int mid_planes = a.pw_bf16 ? 2 * G : Ge;
  unsigned char* wdw = pmid + (size_t)mid_planes * mid_plane;       // [pairs][9][1024]
  unsigned char* wpw = wdw + (size_t)a.pairs * 9 * 1024;            // [cout_p/8][Ge | 2G][8][16]
  const int oy0 = band * a.TH;
  const int th = min(a.TH, a.H - oy0);

  // ---- prologue: model constants only ---------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid < a.n_mt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar_dw[tid])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar_pw[tid])));
  }
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;\n");
  {
    const int4* src = reinterpret_cast<const int4*>(a.wdiag);
    for (int i = tid; i < a.pairs * 9 * 64; i += kThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(wdw) + (uint32_t)i * 16), "l"(src + i));
    // pointwise weights [cout_p][c_p] -> core matrices: item ((g*Ge + kc)*8 + rr), row g*8+rr, chunk kc
    if (!a.pw_bf16) {
      const int items = (a.cout_p / 8) * Ge * 8;
      for (int it = tid; it < items; it += kThreads) {
        const int rr = it & 7, q = it >> 3;
        const int g = q / Ge, kc = q - g * Ge;
        const uint32_t dst = smem_u32(wpw) + (uint32_t)it * 16;
        if (kc < G) {
          const int8_t* src2 = a.pw_w + (size_t)(g * 8 + rr) * a.c_p + kc * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src2));
        } else {
          st_shared16(dst, make_uint4(0, 0, 0, 0));
        }
      }
    } else {                                           // the same integers as bf16, 8 channels per 16 B
      const int items = (a.cout_p / 8) * G * 8;
      for (int it = tid; it < items; it += kThreads) {
        const int rr = it & 7, q = it >> 3;
        const int g = q / G, kc = q - g * G;
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(a.pw_w + (size_t)(g * 8 + rr) * a.c_p + kc * 16));
        const uint32_t dst = smem_u32(wpw) + (uint32_t)(((g * 2 * G + 2 * kc) * 8 + rr) * 16);
        st_shared16(dst, s8x8_to_bf16(w.x, w.y));
        st_shared16(dst + 128, s8x8_to_bf16(w.z, w.w));
      }
    }
  }
  // bias / multiplier vectors: plain loads (a thread stalls on them), so they follow the asynchronous weight copies
  for (int i = tid; i < Ge * 16; i += kThreads) {
    sDwBias[i] = i < a.c_p ? a.dw_bias[i] : 0;
    sDwMult[i] = i < a.c_p ? a.dw_mult[i] : 0.f;
  }
  for (int i = tid; i < a.cout_p; i += kThreads) { sPwBias[i] = a.pw_bias[i]; sPwMult[i] = a.pw_mult[i]; }
  if (Ge > G && !a.pw_bf16)                            // pad group of the depthwise output: zeros
    for (int i = tid; i < a.n_mt * 128; i += kThreads)
      st_shared16(smem_u32(pmid) + (uint32_t)(Ge - 1) * mid_plane + (uint32_t)i * 16, make_uint4(0, 0, 0, 0));
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();

  // ---- fill the input planes ---------------------------------------------------------------
  {
    const uint32_t zpw = (uint32_t)(a.dw_zp_in & 0xff) * 0x01010101u;
    const int n_slots = a.rows_alloc * a.PW;
    for (int i = tid; i < n_slots * G; i += kThreads) {
      const int g = i % G, slot = i / G;               // groups fastest: 16 B x G contiguous in global
      const int ly = (int)__umulhi((uint32_t)slot, a.inv_pw);
      const int lx = slot - ly * a.PW;
      const int iy = oy0 + ly - 1, ix = lx - 1;        // 3x3 SAME, stride 1: pad 1
      const uint32_t dst = smem_u32(pin) + (uint32_t)g * in_plane + (uint32_t)slot * 16;
      if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) {
        st_shared16(dst, make_uint4(zpw, zpw, zpw, zpw));
      } else if (!a.fused_add) {
        const int8_t* src = a.in[0] + (((size_t)b * a.H + iy) * a.W + ix) * a.c_p + g * 16;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
      } else {
        int acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = a.add_init;
        if (a.pre_n) {                                 // inner sum -> its own int8 value -> outer sum
          int acc0[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc0[j] = a.pre_init;
          for (int n = 0; n < 2; ++n) {
            const uint4 v = fetch_resampled(a.in[n], b, iy, ix, a.H, a.W, a.in_h[n], a.in_w[n], a.resample[n],
                                            a.c_p, g * 16);
            const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
            const int mlt = a.pre_mult[n];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc0[q * 4 + 0] += sext_byte<0>(xs[q]) * mlt;
              acc0[q * 4 + 1] += sext_byte<1>(xs[q]) * mlt;
              acc0[q * 4 + 2] += sext_byte<2>(xs[q]) * mlt;
              acc0[q * 4 + 3] += sext_byte<3>(xs[q]) * mlt;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[j] += clampi((acc0[j] >> a.pre_shift) + a.pre_zp, a.pre_lo, a.pre_hi) * a.pre_out_mult;
        }
        for (int n = a.pre_n; n < a.n_in; ++n) {
          const uint4 v = fetch_resampled(a.in[n], b, iy, ix, a.H, a.W, a.in_h[n], a.in_w[n], a.resample[n],
                                          a.c_p, g * 16);
          const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
          const int mlt = a.add_mult[n];
#pragma unroll
          for (int q = 0; q < 4; ++q) {                // sum_i (x_i - zp_i) * m_i = sum_i x_i * m_i + add_init
            acc[q * 4 + 0] += sext_byte<0>(xs[q]) * mlt;
            acc[q * 4 + 1] += sext_byte<1>(xs[q]) * mlt;
            acc[q * 4 + 2] += sext_byte<2>(xs[q]) * mlt;
            acc[q * 4 + 3] += sext_byte<3>(xs[q]) * mlt;
          }
        }
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          packed[q] = vbt::pack4_s8(clampi((acc[q * 4 + 0] >> a.add_shift) + a.add_zp, a.add_lo, a.add_hi),
                                    clampi((acc[q * 4 + 1] >> a.add_shift) + a.add_zp, a.add_lo, a.add_hi),
                                    clampi((acc[q * 4 + 2] >> a.add_shift) + a.add_zp, a.add_lo, a.add_hi),
                                    clampi((acc[q * 4 + 3] >> a.add_shift) + a.add_zp, a.add_lo, a.add_hi));
        st_shared16(dst, make_uint4(packed[0], packed[1], packed[2], packed[3]));
      }
    }
    if (Ge > G)                                        // pad group of the input planes
      for (int i = tid; i < n_slots; i += kThreads)
        st_shared16(smem_u32(pin) + (uint32_t)(Ge - 1) * in_plane + (uint32_t)i * 16, make_uint4(0, 0, 0, 0));
  }
  asm volatile("cp.async.commit_group;\n");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

  // ---- depthwise: nine shifted views per tile and group pair ----------------------------------
  if (warp_u == 0 && elect_one()) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    // descriptors are built once; a start address moves by adding (bytes >> 4) to the low word (everything lies
    // below 256 KB, the 14-bit field never overflows) -- the issuing thread's arithmetic runs on the slow uniform path
    const uint64_t adesc0 = umma_desc(smem_u32(pin), in_plane, 128);
    const uint64_t bdesc0 = umma_desc(smem_u32(wdw), 512, 128);
    for (int mt = 0; mt < a.n_mt; ++mt) {
      for (int p = 0; p < a.pairs; ++p) {
        const uint64_t ad_p = adesc0 + (uint64_t)((((uint32_t)(2 * p) * in_plane) >> 4) + (uint32_t)mt * 128);
        const uint64_t bd_p = bdesc0 + (uint64_t)((uint32_t)(p * 9) * 64);
        const uint32_t d = tmem_u + (uint32_t)(mt * a.pairs + p) * 32;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            umma_i8(d, ad_p + (uint64_t)(uint32_t)(ky * a.PW + kx), bd_p + (uint64_t)((ky * 3 + kx) * 64), idesc,
                    (ky | kx) ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar_dw[mt]));
    }
  }
  // depthwise epilogue: requantise into the middle planes (every position, valid or not)
  for (int mt = 0; mt < a.n_mt; ++mt) {
    block_wait(smem_u32(&mbar_dw[mt]), 0, warp);
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    for (int p = 0; p < a.pairs; ++p) {
      const int grp = 2 * p + half;                    // this warp half requantises one 16-channel group
      if (grp >= G) continue;
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * a.pairs + p) * 32 + half * 16;
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
          "%13, %14, %15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      uint32_t packed[4];
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const int4 bq = *reinterpret_cast<const int4*>(sDwBias + grp * 16 + w4 * 4);
        const float4 mq = *reinterpret_cast<const float4*>(sDwMult + grp * 16 + w4 * 4);
        packed[w4] = a.dw_rq.pack4((int)v[w4 * 4 + 0] + bq.x, (int)v[w4 * 4 + 1] + bq.y, (int)v[w4 * 4 + 2] + bq.z,
                                   (int)v[w4 * 4 + 3] + bq.w, mq.x, mq.y, mq.z, mq.w);
      }
      if (!a.pw_bf16) {
        st_shared16(smem_u32(pmid) + (uint32_t)grp * mid_plane + (uint32_t)(mt * 128 + row) * 16,
                    make_uint4(packed[0], packed[1], packed[2], packed[3]));
      } else {
        const uint32_t dst = smem_u32(pmid) + (uint32_t)(2 * grp) * mid_plane + (uint32_t)(mt * 128 + row) * 16;
        st_shared16(dst, s8x8_to_bf16(packed[0], packed[1]));
        st_shared16(dst + mid_plane, s8x8_to_bf16(packed[2], packed[3]));
      }
    }
  }
  // the middle planes are the pointwise A operand; the depthwise TMEM columns are free again
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");

  // ---- pointwise ---------------------------------------------------------------------------------
  if (warp_u == 0 && elect_one()) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.cout_p >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t abase = smem_u32(pmid), bbase = smem_u32(wpw);
    // bf16: D = F32, A = B = BF16; one MMA (K = 16 elements) per 16-channel group
    const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.cout_p >> 3) << 17) | ((128u >> 4) << 24);
    for (int mt = 0; mt < a.n_mt; ++mt) {
      if (!a.pw_bf16) {
        for (int k2 = 0; k2 < Ge / 2; ++k2)
          umma_i8(tmem_u + (uint32_t)mt * a.cout_p,
                  umma_desc(abase + (uint32_t)(2 * k2) * mid_plane + (uint32_t)mt * 2048, mid_plane, 128),
                  umma_desc(bbase + (uint32_t)k2 * 256, 128, (uint32_t)Ge * 128), idesc, k2 > 0 ? 1u : 0u);
      } else {
        for (int g = 0; g < G; ++g)
          umma_f16(tmem_u + (uint32_t)mt * a.cout_p,
                   umma_desc(abase + (uint32_t)(2 * g) * mid_plane + (uint32_t)mt * 2048, mid_plane, 128),
                   umma_desc(bbase + (uint32_t)g * 256, 128, (uint32_t)(2 * G) * 128), idesc16, g > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar_pw[mt]));
    }
  }
  for (int mt = 0; mt < a.n_mt; ++mt) {
    block_wait(smem_u32(&mbar_pw[mt]), 0, warp);
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const int q = mt * 128 + row;
    const int ly = (int)__umulhi((uint32_t)q, a.inv_pw);
    const int lx = q - ly * a.PW;
    const bool valid = ly < th && lx < a.W;
    int8_t* o = a.out + (size_t)b * a.out_batch_stride + a.out_elem_offset +
                ((size_t)(oy0 + ly) * a.W + lx) * a.out_pix_stride;
    for (int c0 = half * 16; c0 < a.cout_p; c0 += 32) {
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * a.cout_p + c0);
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
          "%13, %14, %15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (!valid) continue;
      if (a.pw_bf16) {                                 // fp32 accumulators hold exact integers
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (uint32_t)__float2int_rn(__uint_as_float(v[j]));
      }
      if (a.vector_out) {
        uint32_t packed[4];
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
          const int4 bq = *reinterpret_cast<const int4*>(sPwBias + c0 + w4 * 4);
          const float4 mq = *reinterpret_cast<const float4*>(sPwMult + c0 + w4 * 4);
          packed[w4] = a.pw_rq.pack4((int)v[w4 * 4 + 0] + bq.x, (int)v[w4 * 4 + 1] + bq.y, (int)v[w4 * 4 + 2] + bq.z,
                                     (int)v[w4 * 4 + 3] + bq.w, mq.x, mq.y, mq.z, mq.w);
        }
        *reinterpret_cast<uint4*>(o + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      } else {                                         // packed head outputs (9 / 36 channels per pixel)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j >= a.cout) break;
          const int y = a.pw_rq((int)v[j] + sPwBias[c0 + j], sPwMult[c0 + j]);
          o[c0 + j] = a.lut ? a.lut[y + 128] : (int8_t)y;
        }
      }
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

// ops[i .. i+n): [ADD] DW PW.  Returns VBT_OK and *taken = true when the fused kernel took them.
int launch_node_umma(const vbt_model* m, const OpRecord* add0, const OpRecord* add, const OpRecord& dw,
                     const OpRecord& pw, const int8_t* const in[3], int8_t* out, long long out_batch_stride, int B,
                     cudaStream_t st, bool* taken) {
  // add0 != null: ADD(add0(in[0], in[1]), in[2]); `pos` = which operand of `add` is add0's output
  *taken = false;
  NodeArgs a;
  a.fused_add = add != nullptr;
  a.pre_n = 0;
  for (int i = 0; i < 3; ++i) {
    a.in[i] = in[i];
    a.in_h[i] = a.in_w[i] = a.resample[i] = a.zp_in[i] = a.add_mult[i] = 0;
  }
  a.pre_mult[0] = a.pre_mult[1] = a.pre_init = a.pre_shift = a.pre_zp = a.pre_lo = a.pre_hi = a.pre_out_mult = 0;
  long long init = add ? (1LL << (add->add_shift > 0 ? add->add_shift - 1 : 0)) : 0;
  if (add && !add0) {
    a.n_in = add->n_in;
    for (int i = 0; i < add->n_in; ++i) {
      a.in_h[i] = add->in_h[i]; a.in_w[i] = add->in_w[i]; a.resample[i] = add->resample[i];
      a.zp_in[i] = add->zp_in[i]; a.add_mult[i] = add->add_mult[i];
      init -= (long long)add->zp_in[i] * add->add_mult[i];
    }
  } else if (add && add0) {
    if (add0->n_in != 2 || add->n_in != 2) return VBT_OK;
    const int pos = add->in[0] == add0->out ? 0 : 1;
    if (add->in[pos] != add0->out || add->resample[pos] != RS_NONE) return VBT_OK;
    a.n_in = 3; a.pre_n = 2;
    long long init0 = 1LL << (add0->add_shift > 0 ? add0->add_shift - 1 : 0);
    for (int i = 0; i < 2; ++i) {
      a.in_h[i] = add0->in_h[i]; a.in_w[i] = add0->in_w[i]; a.resample[i] = add0->resample[i];
      a.pre_mult[i] = add0->add_mult[i];
      init0 -= (long long)add0->zp_in[i] * add0->add_mult[i];
    }
    if (init0 > INT32_MAX || init0 < INT32_MIN) return VBT_OK;
    a.pre_init = (int)init0; a.pre_shift = add0->add_shift; a.pre_zp = add0->zp_out;
    a.pre_lo = add0->act_lo; a.pre_hi = add0->act_hi;
    a.pre_out_mult = add->add_mult[pos];
    init -= (long long)add->zp_in[pos] * add->add_mult[pos];
    const int o = 1 - pos;
    a.in_h[2] = add->in_h[o]; a.in_w[2] = add->in_w[o]; a.resample[2] = add->resample[o];
    a.add_mult[2] = add->add_mult[o];
    init -= (long long)add->zp_in[o] * add->add_mult[o];
  } else {
    a.n_in = 1;
  }
  if (init > INT32_MAX || init < INT32_MIN) return VBT_OK;            // cannot happen for int8 zero points
  a.add_init = (int)init;
  a.add_shift = add ? add->add_shift : 0; a.add_zp = add ? add->zp_out : 0;
  a.add_lo = add ? add->act_lo : 0; a.add_hi = add ? add->act_hi : 0;
  a.B = B; a.H = dw.h_in; a.W = dw.w_in; a.c_p = dw.cout_p;
  a.groups = dw.cout_p / 16; a.pairs = (a.groups + 1) / 2; a.kch_pad = a.pairs * 2;
  a.wdiag = reinterpret_cast<const int8_t*>(m->dev_data + dw.lut_off);
  a.dw_bias = reinterpret_cast<const int32_t*>(m->dev_data + dw.bias_off);
  a.dw_mult = reinterpret_cast<const float*>(m->dev_data + dw.scale_off);
  a.dw_rq = Requant(dw.zp_out, dw.act_lo, dw.act_hi, dw.requant_fast);
  a.dw_zp_in = dw.zp_in[0];
  a.pw_w = reinterpret_cast<const int8_t*>(m->dev_data + pw.w_off);
  a.pw_bias = reinterpret_cast<const int32_t*>(m->dev_data + pw.bias_off);
  a.pw_mult = reinterpret_cast<const float*>(m->dev_data + pw.scale_off);
  a.pw_rq = Requant(pw.zp_out, pw.act_lo, pw.act_hi, pw.requant_fast);
  a.lut = pw.lut_off >= 0 ? reinterpret_cast<const int8_t*>(m->dev_data + pw.lut_off) : nullptr;
  a.cout = pw.cout; a.cout_p = pw.cout_p; a.pw_bf16 = pw.pw_dtype == 1;
  a.out = out; a.out_pix_stride = pw.out_pix_stride; a.out_batch_stride = out_batch_stride;
  a.out_elem_offset = pw.out_elem_offset; a.vector_out = (pw.out_kind == 0);
  // tiling: bands of whole rows, at most kMaxTiles tiles; TMEM = max(depthwise, pointwise) columns
  a.PW = a.W + 2;
  int max_tiles = kMaxTiles;
  while (max_tiles > 1 && std::max(a.pairs * 32, a.cout_p) * max_tiles > 256) --max_tiles;
  if (std::max(a.pairs * 32, a.cout_p) * max_tiles > 512) return VBT_OK;
  // Tile choice: per candidate tile count, estimate waves x per-CTA work (window positions the fill
  // has to produce, halo included, plus a fixed cost for the prologue / MMA / epilogue chain) and
  // keep the cheapest.  A grid a few CTAs over one wave costs a whole extra wave (measured: the
  // 40x40 nodes at 320 CTAs on 296 slots ran 2x longer than at 256).
  static const int cap_tiles = [] { const char* e = getenv("VBT_NODE_MT"); return e ? atoi(e) : kMaxTiles; }();
  static const int fixed_cost = [] { const char* e = getenv("VBT_NODE_FIXED"); return e ? atoi(e) : 400; }();
  max_tiles = std::min(max_tiles, std::max(cap_tiles, 1));
  {
    long long best = -1;
    int best_th = 0, best_mt = 0;
    for (int nt = max_tiles; nt >= 1; --nt) {
      const int th = std::max(1, std::min(a.H, nt * 128 / a.PW));
      const int n_mt = (th * a.PW + 127) / 128;
      if (n_mt > max_tiles) continue;
      const int bands = (a.H + th - 1) / th;
      const int cols_c = std::max(a.pairs * 32, a.cout_p) * n_mt;
      int cols_p = 32;
      while (cols_p < cols_c) cols_p <<= 1;
      const int mid_pl = (pw.pw_dtype == 1) ? 2 * a.groups : a.kch_pad;
      const int plane = std::max(n_mt * 128 + 2 * a.PW + 2 + 8, (th + 2) * a.PW);
      size_t sm = (size_t)a.kch_pad * plane * 16 + (size_t)mid_pl * n_mt * 2048 + (size_t)a.pairs * 9 * 1024 +
                  (size_t)a.cout_p * mid_pl * 16;
      if (sm > 200 * 1024) continue;
      sm = std::max(sm, (size_t)228 * 1024 / (512 / cols_p + 1));
      const int per_sm = std::max(1, std::min((int)(227 * 1024 / (sm + 2304)), 512 / cols_p));
      const long long slots = 148LL * per_sm;
      const long long waves = ((long long)B * bands + slots - 1) / slots;
      const long long cost = waves * ((long long)(th + 2) * a.PW * (add ? 1 + a.n_in : 1) + fixed_cost);
      if (best < 0 || cost < best) { best = cost; best_th = th; best_mt = n_mt; }
    }
    if (best < 0) return VBT_OK;
    a.TH = best_th; a.n_mt = best_mt;
  }
  if (a.n_mt > max_tiles) return VBT_OK;
  a.n_bands = (a.H + a.TH - 1) / a.TH;
  a.rows_alloc = a.TH + 2;
  a.plane_pos = std::max(a.n_mt * 128 + 2 * a.PW + 2 + 8, a.rows_alloc * a.PW);
  int cols = 32;
  while (cols < std::max(a.pairs * 32, a.cout_p) * a.n_mt) cols <<= 1;
  a.tmem_cols = cols;
  a.inv_pw = (uint32_t)((0x100000000ULL + a.PW - 1) / a.PW);
  const int mid_planes = a.pw_bf16 ? 2 * a.groups : a.kch_pad;
  size_t smem = (size_t)a.kch_pad * a.plane_pos * 16 + (size_t)mid_planes * a.n_mt * 2048 +
                (size_t)a.pairs * 9 * 1024 + (size_t)a.cout_p * mid_planes * 16;
  if (smem > 200 * 1024) return VBT_OK;
  smem = std::max(smem, (size_t)228 * 1024 / (512 / cols + 1));
  static bool attr_set = false;
  if (!attr_set) {
    VBT_CHECK_CUDA(cudaFuncSetAttribute(node_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  // VBT_NODE_PRIO=1: launch with the device's greatest priority (measured: no effect on the
  // two-lane pipeline, 31.7 k vs 31.9 k frames/s -- off by default)
  static const bool prio = [] { const char* e = getenv("VBT_NODE_PRIO"); return e && e[0] == '1'; }();
  // VBT_NODE_PDL=0: no programmatic dependent launch for the node kernels (their CTAs then become resident only
  // when the predecessor has finished, instead of holding shared memory / TMEM while they wait)
  static const bool pdl = [] { const char* e = getenv("VBT_NODE_PDL"); return !(e && e[0] == '0'); }();
  VBT_CHECK_CUDA(launch_pdl(node_umma_kernel, dim3((unsigned)a.n_bands, (unsigned)B), dim3(kThreads), smem, st, a, prio, pdl));
  *taken = true;
  return VBT_OK;
}

}  // namespace vbt
