// Pointwise (1x1) convolution, persistent warp-specialised tcgen05 / TMEM int8 GEMM for the
// large-M layers (the 160x160 ... 40x40 maps of the backbone: M = B*H*W up to 1.6 M rows,
// K = cin_p <= 256, N <= 256 per CTA).
//
// replaces: the same CONV_2D 1x1 ops as pw_umma.cu (MBConv expand / project inside
// tflite_runtime's signature_fn(images=...), odt.py:58-61).  pw_umma.cu runs one short-lived
// CTA per 128-row tile; its life is a serial chain (TMEM alloc, weight + bias load, A load,
// MMA, TMEM read-back, store) of ~5 us, and with four CTAs per SM that chain -- not HBM, not
// issue slots -- sets the pace (measured: shrinking the epilogue by a third changed nothing).
// Here a CTA stays resident and walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * prologue (once): TMEM allocation (two accumulators), the weight chunk W[n0:n0+nc, :K]
//     and bias / multipliers into shared memory;
//   * warp 0 (producer): cp.async's the next A tiles into a ring of shared-memory stages, in
//     the K-major no-swizzle core-matrix layout; cp.async.mbarrier.arrive.noinc lets full[s]
//     complete when the copies land, without the producer waiting for them;
//   * warp 1, one thread (MMA): waits full[s] and acc_empty[b], issues K/32 tcgen05.mma.kind::i8
//     (M=128, N=nc) into accumulator b, commits to empty[s] and acc_full[b];
//   * warps 2 .. 2+4E (epilogue, E = 1, 2 or 4 warps per TMEM lane quarter, each taking every E-th
//     16-column chunk): wait acc_full[b], read their 32 TMEM lanes with tcgen05.ld, release the
//     accumulator, requantise (+ residual add), transpose through the quarter's staging buffer
//     and store 32 whole rows with coalesced 128-bit writes.  A single warp per quarter cannot
//     hide the tcgen05.ld / LDS / ALU latencies of its own chain (measured: 2x slower than
//     pw_umma.cu on the wide expand layers), hence E.
// Loads, MMAs and epilogues of consecutive tiles overlap; nothing is re-loaded per tile.
//
// STATUS: default for eligible layers (VBT_PW_PERSIST=0 switches it off).  History, measured on
// B200 (Lite0, frame batch 64; profiles/r1_pw_persist_*.txt): with tiles dealt by a
// fixed stride the kernel LOST to pw_umma.cu (b2.0.expand 105 -> 141 us): cycle counters in the
// three roles showed a resident CTA needs 62 us for its 43 tiles, but not all 296 CTAs become
// resident at once (other kernels' CTAs still hold shared memory / TMEM on some SMs), and with a
// fixed stride the stragglers' tiles waited for a second wave.  With the atomic tile counter a
// late CTA just takes fewer tiles: b1.0.project 80 -> 57 us, b2.1.expand 72 -> 55, b2.0.project
// 45 -> 30, b2.0.expand 105 -> 105; the pipeline gains 2.3 % (36.8 k -> 37.7 k frames/s).
#include "model.cuh"
#include "requant.cuh"

namespace {

using vbt::OpRecord;

constexpr int TILE_M = 128;
constexpr int MAX_NT = 64 + 128 * 4; // producer warp, MMA warp, 4 * E epilogue warps
constexpr int MAX_STAGES = 4;
constexpr int MAX_ACC = 8;           // TMEM accumulators in flight per CTA

struct PwPersistArgs {
  const int8_t* in; const int8_t* res; int8_t* out;
  const int8_t* w; const int32_t* bias; const float* mult;
  long long M;
  int n_tiles;
  int cin_p, cout_p, nc;        // nc: output columns per CTA (multiple of 16, <= 256)
  int kpad;                     // 16-byte K chunks, rounded up to even
  int stages;
  int epi;                      // E: epilogue warps per TMEM lane quarter
  int n_acc;                    // TMEM accumulators (2 .. MAX_ACC): tiles between MMA issue and read-back
  int acc_cols;                 // TMEM columns per accumulator (>= nc)
  int tmem_cols;                // allocation: power of two >= 2 * acc_cols
  int zp_conv, lo, hi;
  int res_zp, add_mult0, add_mult1, add_shift, zp_final;
  int out_stride;               // staging row stride in bytes (odd multiple of 16)
  uint32_t inv_kpad, inv_cpr;   // ceil(65536 / kpad), ceil(65536 / (nc / 16))
  vbt::Requant rq;
  int32_t* counters;            // [16] zero between launches: [0] CTAs done, [1 + y] next tile of N chunk y
};

__device__ __forceinline__ int div_small(int q, uint32_t inv) { return (int)(((uint32_t)q * inv) >> 16); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(n));
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();                  // a lost arrival must not hang the GPU
  }
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <bool HAS_RES>
__global__ void __launch_bounds__(MAX_NT) pw_persist_kernel(PwPersistArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[MAX_STAGES], bar_empty[MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc_full[MAX_ACC], bar_acc_empty[MAX_ACC];
  __shared__ uint32_t tmem_base_s;
  __shared__ int sTile[16];          // tile of the i-th work item of this CTA (ring), -1 = no more work
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NT = (int)blockDim.x;
  const int n0 = blockIdx.y * a.nc;
  const int nc = min(a.nc, a.cout_p - n0);
  const int kpad = a.kpad, kch = a.cin_p >> 4;
  const uint32_t a_bytes = (uint32_t)TILE_M * kpad * 16;
  const uint32_t w_bytes = (uint32_t)a.nc * kpad * 16;
  const uint32_t warp_stage = (uint32_t)32 * a.out_stride;
  unsigned char* sW = smem;
  unsigned char* sA = sW + w_bytes;
  unsigned char* sStage = sA + (size_t)a.stages * a_bytes;            // [4 warps][32 rows][out_stride]
  int32_t* sBias = reinterpret_cast<int32_t*>(sStage + 4 * ((warp_stage + 127) & ~127u));
  float* sMult = reinterpret_cast<float*>(sBias + a.nc);

  // ---- prologue: model constants and this CTA's own state only ------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 32) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 33);           // 32 copy completions + the producer's own arrival
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < a.n_acc; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), 4 * a.epi);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  for (int i = tid; i < nc; i += NT) { sBias[i] = a.bias[n0 + i]; sMult[i] = a.mult[n0 + i]; }
  {
    // weights W[n0 + g*8 + rr][kc*16 ..] -> item ((g * kpad + kc) * 8 + rr), 16 bytes each
    const int items = (nc >> 3) * kpad * 8;
    for (int it = tid; it < items; it += NT) {
      const int rr = it & 7, q = it >> 3;
      const int g = div_small(q, a.inv_kpad), kc = q - g * kpad;
      const bool ok = kc < kch;
      cp_async16(smem_u32(sW) + (uint32_t)it * 16,
                 a.w + (size_t)(n0 + g * 8 + rr) * a.cin_p + (size_t)(ok ? kc : 0) * 16, ok);
    }
    asm volatile("cp.async.commit_group;\n");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  vbt::pdl_wait();                 // activations / residual / output buffer belong to predecessors until here
  vbt::pdl_launch_dependents();

  if (warp == 0) {
    // ---- producer -----------------------------------------------------------------------------
    const int items = TILE_M * kpad;                  // 16-byte granules per A tile (shared: linear)
    // Each lane's cp.async.mbarrier.arrive.noinc makes full[s] count that lane's copies as they
    // land: the producer never blocks on its own loads, so a tile is announced the moment it is
    // in shared memory and the only thing the producer ever waits for is a free stage.
    // Tiles are handed out by an atomic counter, not by a fixed stride: a CTA that becomes resident
    // late (its SM still held CTAs of another kernel, or another stream's) simply takes fewer
    // tiles instead of making the whole launch wait for a second wave.
    int issued = 0;
    for (;; ++issued) {
      const int s = issued % a.stages;
      mbar_wait(smem_u32(&bar_empty[s]), (((uint32_t)(issued / a.stages)) & 1) ^ 1);
      int tile = 0;
      if (lane == 0) tile = atomicAdd(a.counters + 1 + blockIdx.y, 1);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      const bool more = tile < a.n_tiles;
      if (lane == 0) sTile[issued & 15] = more ? tile : -1;
      if (more) {
        const long long m0 = (long long)tile * TILE_M;
        const uint32_t dst0 = smem_u32(sA) + (uint32_t)s * a_bytes;
        for (int it = lane; it < items; it += 32) {
          const int rr = it & 7, q = it >> 3;
          const int g = div_small(q, a.inv_kpad), kc = q - g * kpad;
          const long long m = m0 + g * 8 + rr;
          const bool ok = (m < a.M) && (kc < kch);
          cp_async16(dst0 + (uint32_t)it * 16, a.in + (ok ? m : 0) * a.cin_p + (size_t)(ok ? kc : 0) * 16, ok);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bar_full[s])) : "memory");
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));     // releases sTile[] to the MMA thread
      if (!more) break;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else if (warp == 1) {
    // ---- MMA issuer ------------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nc >> 3) << 17) |
                             ((uint32_t)(TILE_M >> 4) << 24);
      const uint32_t sbo = (uint32_t)kpad * 128;
      for (int it_tile = 0;; ++it_tile) {
        const int s = it_tile % a.stages, b = it_tile % a.n_acc;
        mbar_wait(smem_u32(&bar_acc_empty[b]), (((uint32_t)(it_tile / a.n_acc)) & 1) ^ 1);
        mbar_wait(smem_u32(&bar_full[s]), ((uint32_t)(it_tile / a.stages)) & 1);
        if (sTile[it_tile & 15] < 0) {                // no more work: pass the word on to the epilogue
          mbar_arrive(smem_u32(&bar_acc_full[b]));
          break;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // cp.async (generic proxy) -> MMA reads
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint32_t abase = smem_u32(sA) + (uint32_t)s * a_bytes;
        for (int k2 = 0; k2 < kpad / 2; ++k2)
          umma_i8(tmem + (uint32_t)b * a.acc_cols, umma_desc(abase + k2 * 256, 128, sbo),
                  umma_desc(smem_u32(sW) + k2 * 256, 128, sbo), idesc, k2 > 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[s]));         // stage reusable once these MMAs have read it
        umma_commit(smem_u32(&bar_acc_full[b]));
      }
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes 32 * (w % 4) .. = rows of the tile; the E warps of a
    //      quarter split its columns chunk-wise and share one staging buffer ----------------------
    const int quarter = warp & 3, slice = (warp - 2) >> 2, E = a.epi;
    const int qt = slice * 32 + lane;                 // thread index within the quarter's E warps
    const uint32_t bar_id = 1 + quarter, bar_n = 32 * E;
    unsigned char* stage = sStage + (size_t)quarter * ((warp_stage + 127) & ~127u);
    unsigned char* myrow = stage + (size_t)lane * a.out_stride;
    const int cpr = nc >> 4;
    const uint32_t inv_c = (nc == a.nc) ? a.inv_cpr : (65536u + cpr - 1) / cpr;
    const int round = 1 << (a.add_shift > 0 ? a.add_shift - 1 : 0);
    int last_c0 = -1;                                 // this warp's last chunk (none: nc too narrow)
    for (int c0 = slice * 16; c0 < nc; c0 += 16 * E) last_c0 = c0;
    for (int it_tile = 0;; ++it_tile) {
      const int b = it_tile % a.n_acc;
      mbar_wait(smem_u32(&bar_acc_full[b]), ((uint32_t)(it_tile / a.n_acc)) & 1);
      const int tile = sTile[it_tile & 15];
      if (tile < 0) break;
      const long long mw = (long long)tile * TILE_M + quarter * 32;   // first row of this quarter
      if (HAS_RES) {                                  // residual rows, coalesced, into the staging buffer
        for (int i = qt; i < 32 * cpr; i += 32 * E) {
          const int r = div_small(i, inv_c), j = i - r * cpr;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (mw + r < a.M) v = __ldg(reinterpret_cast<const uint4*>(a.res + (mw + r) * a.cout_p + n0 + j * 16));
          *reinterpret_cast<uint4*>(stage + (size_t)r * a.out_stride + j * 16) = v;
        }
        asm volatile("bar.sync %0, %1;\n" ::"r"(bar_id), "r"(bar_n) : "memory");
      }
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)b * a.acc_cols;
      if (last_c0 < 0) {                              // nothing to read: still hand the accumulator back
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
      }
      for (int c0 = slice * 16; c0 < nc; c0 += 16 * E) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
            "%13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(trow + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c0 == last_c0) {                          // last read of this accumulator: hand it back
          asm volatile("tcgen05.fence::before_thread_sync;\n");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
        }
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (HAS_RES) rv = *reinterpret_cast<const uint4*>(myrow + c0);
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t packed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int4 bq = *reinterpret_cast<const int4*>(sBias + c0 + q * 4);
          const float4 mq = *reinterpret_cast<const float4*>(sMult + c0 + q * 4);
          if (!HAS_RES) {
            packed[q] = a.rq.pack4((int)v[q * 4 + 0] + bq.x, (int)v[q * 4 + 1] + bq.y, (int)v[q * 4 + 2] + bq.z,
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
        *reinterpret_cast<uint4*>(myrow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      asm volatile("bar.sync %0, %1;\n" ::"r"(bar_id), "r"(bar_n) : "memory");
      for (int i = qt; i < 32 * cpr; i += 32 * E) {   // 32 rows x nc bytes: coalesced 128-bit stores
        const int r = div_small(i, inv_c), j = i - r * cpr;
        if (mw + r < a.M)
          *reinterpret_cast<uint4*>(a.out + (mw + r) * a.cout_p + n0 + j * 16) =
              *reinterpret_cast<const uint4*>(stage + (size_t)r * a.out_stride + j * 16);
      }
      asm volatile("bar.sync %0, %1;\n" ::"r"(bar_id), "r"(bar_n) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (tid == 0) {                                     // the last CTA of the launch zeroes the counters again
    __threadfence();
    const int done = atomicAdd(a.counters, 1);
    if (done == (int)(gridDim.x * gridDim.y) - 1) {
      for (int i = 1; i < 16; ++i) a.counters[i] = 0;
      __threadfence();
      a.counters[0] = 0;
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem),
                 "r"((uint32_t)a.tmem_cols));
  }
}

}  // namespace

namespace vbt {

// Returns VBT_OK with *taken = true when the persistent kernel took the op.
int launch_pw_persist(const vbt_model* m, const OpRecord& op, const int8_t* in, const int8_t* res, int8_t* out,
                      int B, cudaStream_t st, bool* taken) {
  *taken = false;
  static const int enabled = [] { const char* e = getenv("VBT_PW_PERSIST"); return e ? atoi(e) : 1; }();
  static const int min_tiles = [] { const char* e = getenv("VBT_PW_PERSIST_MIN_TILES"); return e ? atoi(e) : 592; }();
  if (!enabled || op.out_kind != 0 || op.lut_off >= 0 || !m->cur_counters) return VBT_OK;
  if (op.cin_p % 16 || op.cout_p % 16 || op.cout_p < 16 || op.cin_p > 256) return VBT_OK;
  PwPersistArgs a;
  a.M = (long long)B * op.h_in * op.w_in;
  a.n_tiles = (int)((a.M + TILE_M - 1) / TILE_M);
  if (a.n_tiles < min_tiles) return VBT_OK;
  a.in = in; a.res = res; a.out = out;
  a.counters = m->cur_counters;
  a.w = reinterpret_cast<const int8_t*>(m->dev_data + op.w_off);
  a.bias = reinterpret_cast<const int32_t*>(m->dev_data + op.bias_off);
  a.mult = reinterpret_cast<const float*>(m->dev_data + op.scale_off);
  a.cin_p = op.cin_p; a.cout_p = op.cout_p;
  const int n_chunks = (op.cout_p + 255) / 256;
  a.nc = ((op.cout_p + n_chunks - 1) / n_chunks + 15) / 16 * 16;
  const int kch = op.cin_p / 16;
  a.kpad = (kch + 1) & ~1;
  const bool has_res = res != nullptr;
  a.zp_conv = op.zp_out; a.lo = op.act_lo; a.hi = op.act_hi;
  a.res_zp = op.zp_in[1];
  a.add_mult0 = op.add_mult[0]; a.add_mult1 = op.add_mult[1]; a.add_shift = op.add_shift;
  a.zp_final = op.zp_in[2];
  a.rq = has_res ? Requant(op.zp_out, -128, 127) : Requant(op.zp_out, op.act_lo, op.act_hi, op.requant_fast);
  a.out_stride = ((a.nc / 16) | 1) * 16;
  a.acc_cols = a.nc;
  static const int epi_env = [] { const char* e = getenv("VBT_PW_EPI"); return e ? atoi(e) : 0; }();
  a.epi = a.nc >= 64 ? 4 : (a.nc >= 32 ? 2 : 1);
  if (epi_env == 1 || epi_env == 2 || epi_env == 4) a.epi = epi_env;
  const int NT = 64 + 128 * a.epi;
  static const int acc_env = [] { const char* e = getenv("VBT_PW_ACC"); return e ? atoi(e) : 0; }();
  // accumulators: 2 per CTA with two CTAs per SM, or (VBT_PW_ACC=n) n per CTA from all 512 columns
  a.n_acc = 2;
  int cols = 32;
  if (acc_env >= 2) {
    a.n_acc = std::min(std::min(acc_env, MAX_ACC), 512 / a.acc_cols);
    if (a.n_acc < 2) return VBT_OK;
    cols = 512;
  } else {
    while (cols < 2 * a.acc_cols) cols <<= 1;
  }
  a.tmem_cols = cols;
  a.inv_kpad = (65536u + a.kpad - 1) / a.kpad;
  a.inv_cpr = (65536u + a.nc / 16 - 1) / (a.nc / 16);
  const size_t a_bytes = (size_t)TILE_M * a.kpad * 16, w_bytes = (size_t)a.nc * a.kpad * 16;
  const size_t stage_bytes = 4 * (((size_t)32 * a.out_stride + 127) & ~(size_t)127);
  const int ctas_per_sm = 512 / cols >= 2 ? 2 : 1;
  const size_t budget = (size_t)(ctas_per_sm == 2 ? 100 : 200) * 1024;
  const size_t fixed = w_bytes + stage_bytes + (size_t)a.nc * 8 + 256;
  if (fixed + 2 * a_bytes > budget) return VBT_OK;
  a.stages = (int)std::min<size_t>(MAX_STAGES, (budget - fixed) / a_bytes);
  size_t smem = fixed + (size_t)a.stages * a_bytes;
  // no more resident CTAs than TMEM has columns for
  smem = std::max(smem, (size_t)228 * 1024 / (ctas_per_sm + 1));
  static bool attr_set = false;
  static int n_sm = 148;
  if (!attr_set) {
    VBT_CHECK_CUDA(cudaFuncSetAttribute(pw_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VBT_CHECK_CUDA(cudaFuncSetAttribute(pw_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int dev = 0;
    VBT_CHECK_CUDA(cudaGetDevice(&dev));
    VBT_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int per_chunk = std::max(1, std::min(a.n_tiles, n_sm * ctas_per_sm / n_chunks));
  dim3 grid((unsigned)per_chunk, (unsigned)n_chunks);
  static const int use_pdl = [] { const char* e = getenv("VBT_PW_PERSIST_PDL"); return e ? atoi(e) : 1; }();
  if (!use_pdl) {
    if (has_res) pw_persist_kernel<true><<<grid, dim3(NT), smem, st>>>(a);
    else pw_persist_kernel<false><<<grid, dim3(NT), smem, st>>>(a);
    VBT_CHECK_CUDA(cudaPeekAtLastError());
  } else if (has_res) VBT_CHECK_CUDA(launch_pdl(pw_persist_kernel<true>, grid, dim3(NT), smem, st, a));
  else VBT_CHECK_CUDA(launch_pdl(pw_persist_kernel<false>, grid, dim3(NT), smem, st, a));
  *taken = true;
  return VBT_OK;
}

}  // namespace vbt
