// Pointwise (1x1) convolution as a tcgen05 / TMEM int8 GEMM.
//
// replaces: the CONV_2D 1x1 ops of the EfficientDet-Lite graph inside tflite_runtime's
// signature_fn(images=...) (odt.py:58-61) -- MBConv expand / project, BiFPN lateral and
// node convs, head trunk convs (~90 % of the network's MACs).
//
// GEMM view: D[M, N] = A[M, K] * W[N, K]^T with M = B*H*W pixels, K = cin_p, N = cout_p;
// A and W are K-major int8, the accumulator is int32 in TMEM.  One CTA = one 128-row M
// tile x one N chunk (<= 256 columns):
//   1. cp.async the A tile and the W chunk into shared memory in the canonical
//      no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix);
//   2. one elected thread issues tcgen05.mma.kind::i8 (M=128, N=chunk, K=32 per
//      instruction) over the K extent, then tcgen05.commit -> mbarrier;
//   3. all four warps read their 32 TMEM lanes back with tcgen05.ld and run the int8
//      epilogue (bias, fp32 requantisation, optional quantised residual add, clamp),
//      stage the int8 tile in shared memory and write it out with coalesced 128-bit
//      stores.
// K <= 1152 and N <= 1152 in these networks, M is up to millions of rows: the kernel is
// HBM / epilogue bound, never tensor bound, so it favours many small co-resident CTAs
// (several per SM overlap each other's loads, MMAs and epilogues) over a deep pipeline.
#include "model.cuh"
#include "requant.cuh"

namespace {

using vbt::OpRecord;

constexpr int TILE_M = 128;
constexpr int NT = 256;             // threads per CTA: two warps per TMEM lane quarter, each takes
                                   // half of the chunk's columns in the epilogue
constexpr int KCH_STAGE = 16;      // 16-byte K chunks per K iteration (256 bytes of K)

struct PwUmmaArgs {
  const int8_t* in; const int8_t* res; int8_t* out;
  const int8_t* w; const int32_t* bias; const float* mult;
  long long M;
  int cin_p, cout_p, nc;        // nc: output columns per CTA (multiple of 16, <= 256)
  int zp_conv, lo, hi;
  int res_zp, add_mult0, add_mult1, add_shift, zp_final;
  int out_stride;               // staging row stride in bytes (odd multiple of 16)
  int tmem_cols;                // power of two >= max(32, nc)
  uint32_t inv_kpad[2];         // ceil(65536 / kpad) of a full K stage and of the last one
  uint32_t inv_cpr;             // ceil(65536 / (nc / 16)) per N chunk (last chunk: computed in-kernel)
  vbt::Requant rq;              // conv requant; with a residual: to the full int8 range
};

// q / d for q < 4096, d <= 16, inv = ceil(65536 / d)
__device__ __forceinline__ int div_small(int q, uint32_t inv) { return (int)(((uint32_t)q * inv) >> 16); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                       // src-size 0 -> 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(n));
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrices of 8 rows x 16 B,
// `lbo` bytes between the two K chunks of one MMA, `sbo` bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  return d;                                           // layout_type 0 = no swizzle
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1LL << 24)) __trap();                 // a lost commit must not hang the GPU
  }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <bool HAS_RES, bool FAST>
__global__ void __launch_bounds__(NT) pw_umma_kernel(PwUmmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long m0 = (long long)blockIdx.x * TILE_M;
  const int n0 = blockIdx.y * a.nc;
  const int nc = min(a.nc, a.cout_p - n0);            // last chunk may be narrower (still %16)
  const int kch_total = a.cin_p >> 4;
  const int kch_stage = min(kch_total, KCH_STAGE);
  const int kch_pad = (kch_stage + 1) & ~1;           // MMA K = 32 bytes = 2 chunks
  // shared memory carve-up
  const uint32_t a_bytes = (uint32_t)TILE_M * kch_pad * 16;
  const uint32_t b_bytes = (uint32_t)a.nc * kch_pad * 16;
  const uint32_t stage_bytes = (uint32_t)TILE_M * a.out_stride;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;
  unsigned char* sOut = smem;                                           // aliases A/B after the MMAs
  const uint32_t ab = a_bytes + b_bytes;
  const uint32_t after = (ab > stage_bytes ? ab : stage_bytes);
  unsigned char* sRes = smem + ((after + 127) & ~127u);                 // residual tile (has_res)
  int32_t* sBias = reinterpret_cast<int32_t*>(sRes + (HAS_RES ? ((stage_bytes + 127) & ~127u) : 0));
  float* sMult = reinterpret_cast<float*>(sBias + a.nc);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  for (int i = tid; i < nc; i += NT) { sBias[i] = a.bias[n0 + i]; sMult[i] = a.mult[n0 + i]; }
  vbt::pdl_wait();                 // everything above touched only model constants and this CTA's own state
  vbt::pdl_launch_dependents();
  if (HAS_RES) {                                    // residual tile, row-padded like the output
    const int cpr = nc >> 4;
    const uint32_t inv_c = (nc == a.nc) ? a.inv_cpr : (65536u + cpr - 1) / cpr;
    for (int i = tid; i < TILE_M * cpr; i += NT) {
      const int r = div_small(i, inv_c), j = i - r * cpr;
      const long long m = m0 + r;
      const bool ok = m < a.M;
      cp_async16(smem_u32(sRes + (size_t)r * a.out_stride + j * 16),
                 a.res + (ok ? m : 0) * a.cout_p + n0 + j * 16, ok);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: D=S32, A=B=signed int8, both K-major, M=128, N=nc
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nc >> 3) << 17) |
                         ((uint32_t)(TILE_M >> 4) << 24);
  const uint32_t sbo = (uint32_t)kch_pad * 128;       // 8-row group stride
  uint32_t parity = 0;
  for (int kc0 = 0; kc0 < kch_total; kc0 += KCH_STAGE) {
    const int kch = min(KCH_STAGE, kch_total - kc0);
    const int kpad = (kch + 1) & ~1;
    const uint32_t inv = a.inv_kpad[kch == kch_stage ? 0 : 1];
    // item = ((g * kch_pad + kc) * 8 + rr): shared address = item * 16 (linear), global =
    // row (g*8+rr), chunk kc.  Consecutive lanes walk rr fastest: 8 rows x 16 B = one
    // conflict-free 128-byte shared line; 4 chunks of the same row per warp coalesce.
    const int a_items = (TILE_M / 8) * kpad * 8;
    for (int it = tid; it < a_items; it += NT) {
      const int rr = it & 7, q = it >> 3;
      const int g = div_small(q, inv), kc = q - g * kpad;
      const long long m = m0 + g * 8 + rr;
      const bool ok = (m < a.M) && (kc < kch);
      const uint32_t dst = smem_u32(sA) + (uint32_t)((g * kch_pad + kc) * 8 + rr) * 16;
      cp_async16(dst, a.in + (ok ? m : 0) * a.cin_p + (size_t)(kc0 + (ok ? kc : 0)) * 16, ok);
    }
    const int b_items = (nc / 8) * kpad * 8;
    for (int it = tid; it < b_items; it += NT) {
      const int rr = it & 7, q = it >> 3;
      const int g = div_small(q, inv), kc = q - g * kpad;
      const int n = n0 + g * 8 + rr;
      const bool ok = kc < kch;
      const uint32_t dst = smem_u32(sB) + (uint32_t)((g * kch_pad + kc) * 8 + rr) * 16;
      cp_async16(dst, a.w + (size_t)n * a.cin_p + (size_t)(kc0 + (ok ? kc : 0)) * 16, ok);
    }
    asm volatile("cp.async.commit_group;\n");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes -> tensor-core reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      for (int k2 = 0; k2 < kpad / 2; ++k2) {
        const uint64_t ad = umma_desc(smem_u32(sA) + k2 * 256, 128, sbo);
        const uint64_t bd = umma_desc(smem_u32(sB) + k2 * 256, 128, sbo);
        umma_i8(tmem, ad, bd, idesc, (kc0 > 0 || k2 > 0) ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
          smem_u32(&mbar)) : "memory");
    }
    // MMAs done: smem reusable, TMEM readable.  One warp polls the mbarrier; the rest sleep on the
    // hardware barrier instead of spinning through issue slots co-resident CTAs' epilogues need
    if (warp == 0) mbar_wait(smem_u32(&mbar), parity);
    __syncthreads();
    parity ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n");

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. (rows of the tile); warps 0-3 take the
  //      lower half of the chunk's columns, warps 4-7 the upper half ------------------------
  const int row = ((warp & 3) << 5) | (tid & 31);
  const int split = ((nc >> 4) + 1) / 2 * 16;
  const int c_begin = (warp >> 2) ? split : 0, c_end = (warp >> 2) ? nc : split;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned char* orow = sOut + (size_t)row * a.out_stride;
  const unsigned char* rrow = sRes + (size_t)row * a.out_stride;
  const int round = 1 << (a.add_shift > 0 ? a.add_shift - 1 : 0);
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
        "%13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(trow + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (HAS_RES) rv = *reinterpret_cast<const uint4*>(rrow + c0);
    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t packed[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int4 bq = *reinterpret_cast<const int4*>(sBias + c0 + q * 4);
      const float4 mq = *reinterpret_cast<const float4*>(sMult + c0 + q * 4);
      if (!HAS_RES) {
        packed[q] = a.rq.pack4t<FAST>((int)v[q * 4 + 0] + bq.x, (int)v[q * 4 + 1] + bq.y, (int)v[q * 4 + 2] + bq.z,
                                      (int)v[q * 4 + 3] + bq.w, mq.x, mq.y, mq.z, mq.w);
      } else {
        const int bs[4] = {bq.x, bq.y, bq.z, bq.w};
        const float ms[4] = {mq.x, mq.y, mq.z, mq.w};
        int y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[j] = a.rq((int)v[q * 4 + j] + bs[j], ms[j]);
          const int r = (int)(int8_t)(rw[q] >> (8 * j));
          const int s = (y[j] - a.zp_conv) * a.add_mult0 + (r - a.res_zp) * a.add_mult1 + round;
          y[j] = clampi((s >> a.add_shift) + a.zp_final, a.lo, a.hi);
        }
        packed[q] = vbt::pack4_s8(y[0], y[1], y[2], y[3]);
      }
    }
    *reinterpret_cast<uint4*>(orow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  // ---- coalesced copy-out of the staged tile ----------------------------------------------
  {
    const int cpr = nc >> 4;
    const uint32_t inv_c = (nc == a.nc) ? a.inv_cpr : (65536u + cpr - 1) / cpr;
    for (int i = tid; i < TILE_M * cpr; i += NT) {
      const int r = div_small(i, inv_c), j = i - r * cpr;
      const long long m = m0 + r;
      if (m < a.M)
        *reinterpret_cast<uint4*>(a.out + m * a.cout_p + n0 + j * 16) =
            *reinterpret_cast<const uint4*>(sOut + (size_t)r * a.out_stride + j * 16);
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem),
                 "r"((uint32_t)a.tmem_cols));
  }
}

int pw_impl_from_env() {
  const char* e = getenv("VBT_PW_IMPL");              // "dp4a" forces the SIMT kernel (A/B runs)
  return (e && e[0] == 'd') ? 0 : 1;
}

}  // namespace

namespace vbt {

int launch_pw_persist(const vbt_model* m, const OpRecord& op, const int8_t* in, const int8_t* res, int8_t* out,
                      int B, cudaStream_t st, bool* taken);

int launch_pw_umma(const vbt_model* m, const OpRecord& op, const int8_t* in, const int8_t* res, int8_t* out,
                   long long out_batch_stride, int B, cudaStream_t st, bool* taken) {
  static const int impl = pw_impl_from_env();
  *taken = false;
  // the packed head outputs (9 / 36 channels per pixel, LOGISTIC LUT) stay on the SIMT kernel
  if (!impl || op.out_kind != 0 || op.lut_off >= 0) return VBT_OK;
  if (op.cin_p % 16 || op.cout_p % 16 || op.cout_p < 16) return VBT_OK;
  // large-M, small-K layers: the persistent warp-specialised kernel (pw_persist.cu)
  if (int rc = launch_pw_persist(m, op, in, res, out, B, st, taken)) return rc;
  if (*taken) return VBT_OK;
  PwUmmaArgs a;
  a.in = in; a.res = res; a.out = out;
  a.w = reinterpret_cast<const int8_t*>(m->dev_data + op.w_off);
  a.bias = reinterpret_cast<const int32_t*>(m->dev_data + op.bias_off);
  a.mult = reinterpret_cast<const float*>(m->dev_data + op.scale_off);
  a.M = (long long)B * op.h_in * op.w_in;
  a.cin_p = op.cin_p; a.cout_p = op.cout_p;
  // Output columns per CTA: as wide as one MMA allows (256).  Narrower chunks put more CTAs
  // in flight but repeat the per-CTA fixed work (A tile load, TMEM allocation, barriers);
  // measured on B200 (VBT_PW_NC sweep, profiles/): 128-256 beats 32-96 by 5-20 %.
  static const int nc_max = [] { const char* e = getenv("VBT_PW_NC"); int v = e ? atoi(e) : 256;
                                 return v < 16 ? 16 : (v > 256 ? 256 : v / 16 * 16); }();
  int n_chunks = (op.cout_p + nc_max - 1) / nc_max;
  // small grids (the 10x10 stage: 50 row tiles): narrower N chunks until every SM has a CTA
  static const bool fill = [] { const char* e = getenv("VBT_PW_FILL"); return !(e && e[0] == '0'); }();
  {
    const long long tiles = (a.M + TILE_M - 1) / TILE_M;
    while (fill && tiles * n_chunks < 148 && (op.cout_p + n_chunks) / (n_chunks + 1) >= 32) ++n_chunks;
  }
  a.nc = ((op.cout_p + n_chunks - 1) / n_chunks + 15) / 16 * 16;
  const bool has_res = res != nullptr;
  a.zp_conv = op.zp_out; a.lo = op.act_lo; a.hi = op.act_hi;
  a.res_zp = op.zp_in[1];
  a.add_mult0 = op.add_mult[0]; a.add_mult1 = op.add_mult[1]; a.add_shift = op.add_shift;
  a.zp_final = op.zp_in[2];
  // with a residual the conv result is an int8 intermediate: saturate, clamp after the add
  a.rq = has_res ? Requant(op.zp_out, -128, 127) : Requant(op.zp_out, op.act_lo, op.act_hi, op.requant_fast);
  a.out_stride = ((a.nc / 16) | 1) * 16;
  int cols = 32;
  while (cols < a.nc) cols <<= 1;
  a.tmem_cols = cols;
  const int kch_total = op.cin_p / 16;
  const int kch = std::min(kch_total, KCH_STAGE), kpad = (kch + 1) & ~1;
  const int kch_last = kch_total - (kch_total - 1) / KCH_STAGE * KCH_STAGE, kpad_last = (kch_last + 1) & ~1;
  a.inv_kpad[0] = (65536u + kpad - 1) / kpad;
  a.inv_kpad[1] = (65536u + kpad_last - 1) / kpad_last;
  a.inv_cpr = (65536u + a.nc / 16 - 1) / (a.nc / 16);
  const size_t ab = (size_t)(TILE_M + a.nc) * kpad * 16;
  const size_t stage = ((size_t)TILE_M * a.out_stride + 127) & ~(size_t)127;
  size_t smem = ((std::max(ab, stage) + 127) & ~(size_t)127) + (has_res ? stage : 0) + (size_t)a.nc * 8 + 128;
  // TMEM holds 512 columns per SM: ask for enough shared memory that no more CTAs become
  // resident than can own `cols` columns each, so tcgen05.alloc never has to spin
  const size_t cap_ctas = 512 / cols;
  smem = std::max(smem, (size_t)228 * 1024 / (cap_ctas + 1));
  static bool attr_set = false;
  if (!attr_set) {
    VBT_CHECK_CUDA(cudaFuncSetAttribute(pw_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VBT_CHECK_CUDA(cudaFuncSetAttribute(pw_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VBT_CHECK_CUDA(cudaFuncSetAttribute(pw_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (smem > 200 * 1024) return VBT_OK;
  dim3 grid((unsigned)((a.M + TILE_M - 1) / TILE_M), (unsigned)((op.cout_p + a.nc - 1) / a.nc));
  if (has_res) VBT_CHECK_CUDA(launch_pdl(pw_umma_kernel<true, false>, grid, dim3(NT), smem, st, a));
  else if (a.rq.fast) VBT_CHECK_CUDA(launch_pdl(pw_umma_kernel<false, true>, grid, dim3(NT), smem, st, a));
  else VBT_CHECK_CUDA(launch_pdl(pw_umma_kernel<false, false>, grid, dim3(NT), smem, st, a));
  *taken = true;
  return VBT_OK;
}

}  // namespace vbt
