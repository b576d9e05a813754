// Fused MBConv block:  1x1 expand (ReLU6) -> depthwise 3x3 / 5x5, stride 1 / 2 (ReLU6) -> 1x1 project
// [+ residual], ONE kernel per block; the 6x expanded tensor and the depthwise output never reach HBM.
//
// replaces: the CONV_2D 1x1 -> DEPTHWISE_CONV_2D -> CONV_2D 1x1 [-> ADD] quadruples of the
// EfficientNet-Lite backbone inside tflite_runtime's signature_fn(images=...) (odt.py:58-61); SURVEY.md
// 2.4 K3 / appendix A.2: 16 / 21 / 21 blocks in Lite0 / 1 / 2, 86 % of the network's MACs and (unfused)
// 27.9 of its 35 M activation elements per frame.
//
// One CTA = one frame x one TH x TW tile of OUTPUT pixels, all channels.  The expanded channels are
// walked in chunks of 32 (one tcgen05 N = 32 column block = two 16-channel groups):
//   fill     the tile's input window (WH x WW pixels, halo included) -> shared-memory planes, one per
//            16-channel group: [group][m][16 B], m = wy * WW + wx: the K-major core-matrix A operand
//            of the expand GEMM.
//   per chunk c (weights: ONE bulk copy of a host-prepared image, four buffers deep, requested two chunks ahead):
//     E   expand GEMM  D_E[window positions, 32] = bias + in planes x Wexp_c^T   (tcgen05.mma kind::i8,
//         accumulators in TMEM, pre-loaded with the chunk's bias by tcgen05.st), issued one chunk ahead so
//         that it runs under the previous chunk's depthwise
//     EE  epilogue: requantise + ReLU6 -> the chunk's expanded tensor as CHANNEL-PLANAR UNSIGNED bytes
//         (value + 128) [window row][channel][cs bytes]; positions outside the image get the expanded tensor's
//         zero point (TF SAME pads the depthwise INPUT, i.e. the expanded tensor).  Per value: I2F, FMUL, half
//         a FADD2, one VIADDMNMX.RELU, one byte store at an immediate offset (ee_tiles)
//     DW  depthwise on the SIMT pipes: lane = channel, a warp walks strips of four output pixels of
//         one row.  In the planar layout four horizontally adjacent taps of ONE channel are the four
//         bytes of a word, so a 5-tap row of the window is two dp4a.u32.s32 (3-tap: one) with all lanes
//         useful -- no masked weights, no unpacking (the bias carries - 128 * sum(w)); windows at odd byte
//         offsets come from PRMT.  Weights (K rows x 1 or 2 words), bias and multiplier of the lane's channel
//         stay in registers for the chunk.  Requantise + ReLU6 -> two "middle" planes [group][q][16 B] = the K-major A operand of
//         the project GEMM.  (Two other forms were built first and measured on B200: the tensor-pipe
//         depthwise of csrc/dw_umma.cu -- block-diagonal tap matrices -- occupies the pipe ~100
//         cycles per M128 N32 K32 tap MMA, 2,500 cycles per chunk for 5x5, on the pipe the two GEMMs
//         need; position-major dp4a against pre-masked weight words wastes 3 of 4 lanes: 3,300.)
//     P   project GEMM  D_P[output positions, Cout] += middle planes x Wproj[:, chunk]^T, the
//         accumulator staying in TMEM over all chunks
//   final   D_P -> bias, requantise, quantised residual add (the block input, still in the input
//           planes), store.
// MMAs are issued by one ELECTED lane of a warp-uniform region: issuing from `if (tid == 0)` makes
// ptxas wrap every UTCIMMA in an ELECT / R2UR / BRA.U.ANY loop (~100 cycles each, measured).
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <string.h>

#include <utility>

#include "model.cuh"
#include "requant.cuh"

namespace {

using vbt::OpRecord;

constexpr int kThreads = 256;      // 8 warps: warp w owns TMEM lanes 32 * (w % 4) .., 16-channel half w / 4
constexpr int kWBuf = 4;           // weight-image buffers: images are requested three chunks ahead
constexpr int kMaxCout = 352;

struct MbArgs {
  const int8_t* in; int8_t* out;
  const unsigned char* img;                  // [n_chunks][img_stride] weight images (effdet.mbconv_images)
  const int32_t* pj_bias; const float* pj_mult;
  int B, H, W, Ho, Wo;                       // depthwise input (= block input) size, output size
  int cin_p, g_in, ge_in, cout_p, n_chunks;
  int pad_top, pad_left;
  int has_expand, has_res;
  int stem, img_h, img_w, stem_pad_top, stem_pad_left, stem_zp;   // expand stage = the stem conv over im2col rows of the uint8 frame
  // tile geometry: TH x TW output pixels; window WH x WW input pixels, M index m = wy * WW + wx
  int TH, TW, TWp, tiles_x, WH, WW, m_total, n_win_tiles, n_out_tiles, strips_x, n_strips;
  int row_stride, chan_stride;               // planar expanded planes: bytes per window row / per channel
  uint32_t inv_ww, inv_twp, inv_gin, inv_sx; // ceil(2^32 / d)
  int zp_fill;                               // zero point of the depthwise input (padding value)
  int ex_off, ex_span, ex_ulo;               // expand epilogue, packed clamp: relu(min(bits + ex_off, ex_span)) + ex_ulo = value + 128
  vbt::Requant ex_rq, dw_rq, pj_rq;
  int pj_zp, res_zp, add_mult0, add_mult1, add_shift, zp_final, lo, hi;
  int img_stride, img_bytes, off_taps, off_wproj, off_consts;
  int tmem_cols, col_pj;
  int use_tma;                               // the input window arrives as tiled TMA loads (one per 16-channel group)
  int split, col_split, stage_stride;          // split 2: a cluster of two CTAs shares a tile's chunks (see kernel)
  uint32_t sm_stage;
  uint32_t sm_exp, sm_mid, sm_wbuf, in_gstride, mid_gstride;   // bytes
  long long* dbg;                            // VBT_MB_DBG=1: cycle counters of CTA (0,0), threads 0 and 64
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
// Warp-uniform issue: the whole warp runs the (uniform) descriptor arithmetic, one elected lane issues.
// Issuing from a divergent `if (tid == 0)` makes ptxas wrap every UTCIMMA in an ELECT / BRA.U.ANY loop
// (its operands live in uniform registers and it cannot prove a one-lane region uniform): ~100 cycles
// per instruction, measured.
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
    if (spin > (1LL << 24)) __trap();                 // a lost commit must not hang the GPU
  }
}
__device__ __forceinline__ void st_shared16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ uint4 ld_shared16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }


__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster16(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// (compiled in only for the DBG instantiations: even switched off at run time, two extra tick sites cost 5 % --
// registers and code layout -- measured on B200)
#define MB_TICK(i)                                                        \
  do {                                                                   \
    if (DBG && dbg_on) { const long long now__ = clock64(); dbg_acc[i] += now__ - dbg_t; dbg_t = now__; } \
  } while (0)

// 16 int32 / float constants of this thread's channel half from the chunk image
__device__ __forceinline__ void load16(uint32_t addr, int (&o)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = ld_shared16(addr + j * 16);
    o[j * 4] = (int)v.x; o[j * 4 + 1] = (int)v.y; o[j * 4 + 2] = (int)v.z; o[j * 4 + 3] = (int)v.w;
  }
}
__device__ __forceinline__ void load16(uint32_t addr, float (&o)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 v = ld_shared16(addr + j * 16);
    o[j * 4] = __uint_as_float(v.x); o[j * 4 + 1] = __uint_as_float(v.y);
    o[j * 4 + 2] = __uint_as_float(v.z); o[j * 4 + 3] = __uint_as_float(v.w);
  }
}
__device__ __forceinline__ uint32_t ld_shared32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const int (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// The expanded tensor lives in shared memory as bytes [window row][channel 0..31][cs bytes of that row]:
// four horizontally adjacent positions of ONE channel are a word (what the depthwise wants), and the 16
// channels of a position lie a fixed `cs` apart.  The bytes are stored UNSIGNED (value + 128): the expand
// epilogue's clamp then ends in [0, hi - lo] with nothing left to add, and the depthwise multiplies with
// dp4a.u32.s32 against a bias that carries - 128 * sum(w) (effdet.mbconv_images).
// 16 values (low bytes of 16 registers) of one position -> its 16 channel rows.  cs is a template parameter
// so that the 16 stores share one address register (cs = 8 k + 4: an odd number of words, so that lanes =
// channels of the depthwise hit 32 different banks).
// CS > 0: compile-time channel stride (the 16 stores share one address register: the offsets are immediates);
// CS == 0: `cs` at run time
template <int OFF>
__device__ __forceinline__ void st_shared8_imm(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0+%2], %1;\n" ::"r"(addr), "r"(v), "n"(OFF));
}
template <int CS, int... J>
__device__ __forceinline__ void scatter16_imm(uint32_t dst, const uint32_t (&u)[16], std::integer_sequence<int, J...>) {
  (st_shared8_imm<J * CS>(dst, u[J]), ...);
}
template <int CS>
__device__ __forceinline__ void scatter16(uint32_t dst, uint32_t cs, const uint32_t (&u)[16]) {
  if (CS) {
    scatter16_imm<CS>(dst, u, std::make_integer_sequence<int, 16>());
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) st_shared8(dst + (uint32_t)j * cs, u[j]);
  }
}
struct EeCtx {
  uint32_t tmem;                 // this thread's TMEM row, first of its 16 columns
  uint32_t dst;                  // its 16 channels' rows in the expanded planes (shared-memory address)
  const uint32_t* pos;           // &sPos[0][row]
  int n_win, wt0, wt_step;
  uint32_t window_mask, inside_mask, zpu;
  int off, span, ulo, fast, store_next;
};
// Expand epilogue of one chunk: accumulators (bias included) of this thread's window positions -> requantise,
// ReLU6 -> unsigned bytes in the 16 channel rows.  Fast path per value: I2F, FMUL, half a FADD2 (round by magic
// add: the integer lands in the low mantissa bits), ONE add-min-relu on those bits =
// clamp(round, lo - zp, hi - zp) - (lo - zp), which IS the stored byte when lo = -128, and the byte store.
// Positions outside the image take the zero point (TF SAME pads the depthwise INPUT): a select after the
// arithmetic, entered only by warps that hold such a position -- no divergent second path.
template <int CS>
__device__ __forceinline__ void ee_tiles(const EeCtx& e, const vbt::Requant& rq, const float (&em)[16], const int (&eb_next)[16],
                                         uint32_t cs = 0) {
  for (int wt = e.wt0; wt < e.n_win; wt += e.wt_step) {
    uint32_t v[16];
    tmem_ld16(e.tmem + (uint32_t)(wt * 32), v);
    // the accumulators are in registers: the next chunk's bias takes their place at once, so that the store's
    // latency hides under this tile's arithmetic (completion is awaited once, after the last tile)
    if (e.store_next) tmem_st16(e.tmem + (uint32_t)(wt * 32), eb_next);
    const bool inside = (e.inside_mask >> wt) & 1u;
    const bool all_inside = __all_sync(0xffffffffu, inside);
    const uint32_t dst = e.dst + e.pos[wt * 128];
    uint32_t u[16];
    if (e.fast) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float p0 = __fmul_rn(__int2float_rn((int)v[j]), em[j]), p1 = __fmul_rn(__int2float_rn((int)v[j + 1]), em[j + 1]);
        unsigned long long q, mg;
        uint32_t b0, b1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(p0), "f"(p1));
        asm("mov.b64 %0, {%1, %1};" : "=l"(mg) : "f"(vbt::kRoundMagic));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(mg));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(b0), "=r"(b1) : "l"(q));
        u[j] = (uint32_t)__viaddmin_s32_relu((int)b0, e.off, e.span);
        u[j + 1] = (uint32_t)__viaddmin_s32_relu((int)b1, e.off, e.span);
      }
      if (e.ulo) {
#pragma unroll
        for (int j = 0; j < 16; ++j) u[j] += (uint32_t)e.ulo;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) u[j] = (uint32_t)(rq((int)v[j], em[j]) + 128);
    }
    if (!all_inside) {
#pragma unroll
      for (int j = 0; j < 16; ++j) u[j] = inside ? u[j] : e.zpu;
    }
    if ((e.window_mask >> wt) & 1u) scatter16<CS>(dst, cs, u);
  }
}
// dp4a with unsigned activation bytes and signed weight bytes
__device__ __forceinline__ int dp4a_us(uint32_t x, uint32_t w, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;\n" : "=r"(d) : "r"(x), "r"(w), "r"(acc));
  return d;
}
// bytes OFF .. OFF + 3 of the byte string (w0, w1, w2)
template <int OFF>
__device__ __forceinline__ uint32_t window(uint32_t w0, uint32_t w1, uint32_t w2) {
  constexpr int o = OFF & 3;
  constexpr uint32_t sel = (uint32_t)(o | ((o + 1) << 4) | ((o + 2) << 8) | ((o + 3) << 12));
  if (OFF == 0) return w0;
  if (OFF == 4) return w1;
  if (OFF == 8) return w2;
  return OFF < 4 ? __byte_perm(w0, w1, sel) : __byte_perm(w1, w2, sel);
}
// byte IDX of (w0, w1, w2) in the low byte (the other bytes meet zero weights)
template <int IDX>
__device__ __forceinline__ uint32_t byte_at(uint32_t w0, uint32_t w1, uint32_t w2) {
  const uint32_t w = IDX < 4 ? w0 : (IDX < 8 ? w1 : w2);
  return (IDX & 3) ? (w >> (8 * (IDX & 3))) : w;
}

template <int K, int S, int MINB, bool DBG>
__global__ void __launch_bounds__(MINB == 1 ? 512 : 256, MINB) mbconv_umma_kernel(MbArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_w[kWBuf], bar_e, bar_p[2], bar_in;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) int32_t sPjBias[kMaxCout];
  __shared__ __align__(16) float sPjMult[kMaxCout];
  __shared__ uint32_t sPos[6][128];                   // byte offset of window position (tile, row) inside a channel's rows
  constexpr int NT = MINB == 1 ? 512 : 256;           // MINB == 1: sixteen warps, for tiles that leave room for one CTA per SM only
  constexpr int NWW = K == 5 ? 2 : 1;                 // weight words per window row
  constexpr int NLD = S == 1 ? 2 : 3;                 // activation words per window row and strip
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, half = (tid >> 7) & 1, tgrp = tid >> 8;     // tgrp: which tiles this half of a 512-thread CTA takes
  // split 2: the two CTAs of a cluster work on the same tile, CTA r on the chunks r, r + 2, ... ; their
  // partial project accumulators meet through distributed shared memory at the end
  const int split = a.split;
  const int rank = split == 2 ? (int)cluster_ctarank() : 0;
  const int tile = blockIdx.x / split, b = blockIdx.y;
  const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
  const int oy0 = ty * a.TH, ox0 = tx * a.TW;
  const int ey0 = oy0 * S - a.pad_top, ex0 = ox0 * S - a.pad_left;   // window origin in input coordinates
  const uint32_t s_in = smem_u32(smem), s_exp = s_in + a.sm_exp, s_mid = s_in + a.sm_mid;
  const uint32_t s_wbuf = s_in + a.sm_wbuf;
  const int n_chunks = (a.n_chunks - rank + split - 1) / split;     // this CTA's chunks (local index c)
  const uint32_t cs = (uint32_t)a.chan_stride;
  const bool dbg_on = DBG && a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 64);
  long long dbg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long dbg_t = dbg_on ? clock64() : 0;

  // ---- prologue: model constants and CTA-private state only --------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  auto load_image = [&](int c) {   // one thread: one bulk copy (TMA engine) of chunk c's weight image
    const uint32_t bar = smem_u32(&bar_w[c % kWBuf]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)a.img_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(s_wbuf + (uint32_t)(c % kWBuf) * a.img_stride),
                 "l"(a.img + (size_t)(rank + c * split) * a.img_stride), "r"((uint32_t)a.img_bytes), "r"(bar)
                 : "memory");
  };
  // The threads that issue the first bulk copies initialise the barriers those copies signal themselves and issue
  // at once: the copies' round trips (weights: model constants; input window: after the PDL wait) run under the
  // TMEM allocation and the block barrier instead of behind them.
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_e)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_p[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_p[1])));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  if (tid == 64) {
    for (int i = 0; i < kWBuf; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_w[i])));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    load_image(0);
    if (n_chunks > 1) load_image(1);
    if (n_chunks > 2) load_image(2);
  }
  if (tid == 96) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_in)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (a.has_expand && a.use_tma && !a.stem) {
      // -> position-major planes, the expand GEMM's A operand, by TMA: the activation tensor is described to
      // the copy engine as [B][H][W][cin_p] bytes; one tiled load per 16-channel group drops the group's
      // WH x WW window (16-byte rows) exactly where the plane layout wants it -- [m = wy * WW + wx][16 B] --
      // and fills what lies outside the image with zeros (never used: the epilogue overrides those positions).
      vbt::pdl_wait();
      const uint32_t bar = smem_u32(&bar_in);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(a.g_in * a.m_total * 16)) : "memory");
      for (int g = 0; g < a.g_in; ++g)
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(s_in + (uint32_t)g * a.in_gstride), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(g * 16), "r"(ex0), "r"(ey0), "r"(b), "r"(bar)
            : "memory");
    }
  }
  for (int i = tid; i < a.cout_p; i += NT) { sPjBias[i] = a.pj_bias[i]; sPjMult[i] = a.pj_mult[i]; }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();                 // barriers initialised before anyone arms or polls them; TMEM base published
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();

  // ---- fill: the tile's input window ------------------------------------------------------------------
  const uint32_t zpu = (uint32_t)((a.zp_fill + 128) & 0xff);     // the padding value as stored (unsigned)
  if (a.stem) {
    // the stem as the expand stage: window position (wy, wx) is a pixel of the stem's OUTPUT; its GEMM row
    // is the 3x3x3 patch of the uint8 frame around (2 sy, 2 sx), bytes k = (ky * 3 + kx) * 3 + c in
    // 0 .. 26 (27 .. 31 zero), split over the two 16-byte K planes.  One thread = one position.
    const uint8_t* fin = reinterpret_cast<const uint8_t*>(a.in) + (size_t)b * a.img_h * a.img_w * 3;
    const uint32_t zp = (uint32_t)a.stem_zp;
    for (int m = tid; m < a.m_total; m += NT) {
      const int wy = (int)__umulhi((uint32_t)m, a.inv_ww);
      const int wx = m - wy * a.WW;
      const int sy = ey0 + wy, sx = ex0 + wx;
      if (sy < 0 || sy >= a.H || sx < 0 || sx >= a.W) continue;          // EE overrides these positions
      const int iy0 = 2 * sy - a.stem_pad_top, ix0 = 2 * sx - a.stem_pad_left;
      const uint32_t d0 = s_in + (uint32_t)m * 16, d1 = d0 + a.in_gstride;
      // interior: the nine bytes of a patch row are contiguous in the frame -> three aligned words + one
      // funnel shift per word; the 27 bytes are then permuted into the two 16-byte K planes in registers.
      // (The aligned window may start up to 3 bytes early and end up to 3 bytes late: kept inside the
      // frame buffer by sending the frame's last rows to the byte-wise path.)
      if (iy0 >= 0 && iy0 + 3 < a.img_h && ix0 >= 0 && ix0 + 3 <= a.img_w) {
        uint32_t A[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint8_t* p = fin + ((size_t)(iy0 + ky) * a.img_w + ix0) * 3;
          const uint32_t sh = ((uint32_t)(uintptr_t)p & 3u) * 8;
          const uint32_t* q = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
          const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
          A[ky][0] = __funnelshift_r(w0, w1, sh);          // bytes 0..3 of the row
          A[ky][1] = __funnelshift_r(w1, w2, sh);          // bytes 4..7
          A[ky][2] = __funnelshift_r(w2, 0u, sh);          // byte 8 (+ junk)
        }
        // K bytes: row0[0..8] row1[0..8] row2[0..8] 0 0 0 0 0
        const uint32_t k0 = A[0][0], k1 = A[0][1];
        const uint32_t k2 = __byte_perm(A[0][2], A[1][0], 0x6540);                       // r0b8 r1b0 r1b1 r1b2
        const uint32_t k3 = __byte_perm(A[1][0], A[1][1], 0x6543);                       // r1b3 r1b4 r1b5 r1b6
        const uint32_t k4 = __byte_perm(__byte_perm(A[1][1], A[1][2], 0x0043), A[2][0], 0x5410);   // r1b7 r1b8 r2b0 r2b1
        const uint32_t k5 = __byte_perm(A[2][0], A[2][1], 0x5432);                       // r2b2 .. r2b5
        const uint32_t k6 = __byte_perm(A[2][1], A[2][2], 0x0432) & 0x00ffffffu;         // r2b6 r2b7 r2b8 0
        st_shared16(d0, make_uint4(k0, k1, k2, k3));
        st_shared16(d1, make_uint4(k4, k5, k6, 0u));
        continue;
      }
      // frame border: byte by byte, taps outside the frame carry the input zero point (TF SAME)
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = iy0 + ky;
        const bool row_ok = iy >= 0 && iy < a.img_h;
        const uint8_t* src = fin + ((size_t)(row_ok ? iy : 0) * a.img_w + ix0) * 3;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int ix = ix0 + j / 3;
          uint32_t v = zp;
          if (row_ok && ix >= 0 && ix < a.img_w) v = __ldg(src + j);
          const int k = ky * 9 + j;
          st_shared8((k >> 4 ? d1 : d0) + (uint32_t)(k & 15), v);
        }
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) st_shared8(d1 + (uint32_t)(k & 15), 0u);
    }
  } else if (a.has_expand && a.use_tma) {
    // the window arrives by TMA, issued in the prologue (thread 96)
  } else if (a.has_expand) {       // the same by per-thread cp.async (VBT_MB_TMA=0, or no tensor map could be encoded)
    const int G = a.g_in;
    const int8_t* fin = a.in + (size_t)b * a.H * a.W * a.cin_p;
    const int items = a.m_total * G;
    for (int i = tid; i < items; i += NT) {
      const int m = G == 1 ? i : (int)__umulhi((uint32_t)i, a.inv_gin);   // ceil(2^32 / 1) does not fit
      const int g = i - m * G;                          // groups fastest: 16 B x G contiguous in global
      const int wy = (int)__umulhi((uint32_t)m, a.inv_ww);
      const int wx = m - wy * a.WW;
      const int iy = ey0 + wy, ix = ex0 + wx;
      if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {  // positions outside the image are never used: EE overrides them
        const int8_t* src = fin + ((size_t)iy * a.W + ix) * a.cin_p + g * 16;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s_in + (uint32_t)g * a.in_gstride + (uint32_t)m * 16), "l"(src));
      }
    }
  } else {                         // no expand conv: the input IS the depthwise input -> channel planes
    const int8_t* fin = a.in + (size_t)b * a.H * a.W * a.cin_p;
    for (int i = tid; i < a.m_total * 2; i += NT) {
      const int m = i >> 1, g = i & 1;
      const int wy = (int)__umulhi((uint32_t)m, a.inv_ww);
      const int wx = m - wy * a.WW;
      const int iy = ey0 + wy, ix = ex0 + wx;
      uint32_t u[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) u[j] = zpu;
      if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W && g < a.g_in) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)iy * a.W + ix) * a.cin_p + g * 16));
        const uint32_t w[4] = {v.x ^ 0x80808080u, v.y ^ 0x80808080u, v.z ^ 0x80808080u, v.w ^ 0x80808080u};
#pragma unroll
        for (int j = 0; j < 16; ++j) u[j] = w[j >> 2] >> (8 * (j & 3));
      }
      scatter16<0>(s_exp + (uint32_t)(g * 16) * cs + (uint32_t)(wy * a.row_stride + wx), cs, u);
    }
  }
  // which of this thread's window positions (row of every window tile) lie inside the image / the window
  uint32_t inside_mask = 0, window_mask = 0;
  for (int wt = 0; wt < a.n_win_tiles; ++wt) {
    const int m = wt * 128 + row;
    const int wy = (int)__umulhi((uint32_t)m, a.inv_ww);
    const int wx = m - wy * a.WW;
    const int iy = ey0 + wy, ix = ex0 + wx;
    if (m < a.m_total) window_mask |= 1u << wt;
    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) inside_mask |= 1u << wt;
    if (half == 0 && tgrp == 0) sPos[wt][row] = (uint32_t)(wy * a.row_stride + wx);
  }
  // the expand accumulators start from the bias: every thread stores its 16 channels' bias of chunk 0 into its
  // TMEM row of every window tile, and the expand MMAs always accumulate
  if (a.has_expand) {
    mbar_wait(smem_u32(&bar_w[0]), 0);
    int eb[16];
    load16(s_wbuf + a.off_consts + half * 64, eb);
    for (int wt = tgrp; wt < a.n_win_tiles; wt += NT / 256) tmem_st16(tmem + lane_base + (uint32_t)(wt * 32 + half * 16), eb);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    if (a.use_tma && !a.stem) mbar_wait(smem_u32(&bar_in), 0);
  }
  asm volatile("cp.async.commit_group;\n");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // the issuing warps see their own index and the TMEM base as warp-uniform values
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  MB_TICK(0);                                          // prologue + fill

  // D = S32, A = B = signed int8, K-major, M = 128, N = 32
  const uint32_t idesc32 = (2u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_e = a.stem ? (idesc32 & ~(7u << 7)) : idesc32;      // the stem's A operand is UNSIGNED int8 (format 0)
  // Descriptors are built once; inside the loops a start address moves by adding (bytes >> 4) to the
  // descriptor's low word (the 14-bit address field never overflows: everything lies below 256 KB).
  // Two issuing warps: warp 0 owns P, warp 1 owns E, one elected lane each.
  const uint64_t e_adesc = umma_desc(s_in, a.in_gstride, 128);
  const uint64_t e_bdesc = umma_desc(s_wbuf, 128, (uint32_t)a.ge_in * 128);
  const uint32_t e_kstep = (2 * a.in_gstride) >> 4;
  auto issue_expand = [&](int c) {                     // one elected lane of warp 1
    const uint64_t bd = e_bdesc + (uint64_t)(((uint32_t)(c % kWBuf) * a.img_stride) >> 4);
    const int ksteps = a.ge_in >> 1;
    for (int wt = 0; wt < a.n_win_tiles; ++wt) {
      const uint64_t ad = e_adesc + (uint64_t)(wt * 128);
      const uint32_t d = tmem_u + (uint32_t)wt * 32;
      for (int k2 = 0; k2 < ksteps; ++k2)
        umma_i8(d, ad + (uint64_t)(k2 * e_kstep), bd + (uint64_t)(k2 * 16), idesc_e, 1u);   // onto the bias
    }
    umma_commit(smem_u32(&bar_e));
  };
  if (warp_u == 1 && a.has_expand) {       // image 0: every thread polled its barrier before storing the bias
    if (elect_one()) issue_expand(0);
  }

  const uint64_t pj_adesc = umma_desc(s_mid, a.mid_gstride, 128);
  const uint64_t pj_bdesc = umma_desc(s_wbuf + a.off_wproj, 128, 256);
  const int pj_n0 = a.cout_p <= 256 ? a.cout_p : (a.cout_p / 32) * 16;
  const uint32_t pj_id0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(pj_n0 >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t pj_id1 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((a.cout_p - pj_n0) >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t mid_buf = 2 * a.mid_gstride;          // bytes of one middle-plane buffer (two groups)
  uint32_t par_e = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t wb = s_wbuf + (uint32_t)(c % kWBuf) * a.img_stride;
    const uint32_t consts = wb + a.off_consts;
    // images c and c + 1 (its expand bias is stored at the end of this chunk's epilogue): ONE warp polls the
    // barriers, the block barrier below hands the visibility on (image c + 1 was requested a chunk ago)
    if (!a.has_expand) mbar_wait(smem_u32(&bar_w[c % kWBuf]), (uint32_t)((c / kWBuf) & 1));
    MB_TICK(1);
    // ---- EE: expand epilogue -> the chunk's expanded tensor, channel-planar ----------------------------
    if (a.has_expand) {
      // E(c) complete and image c + 1 landed: for c >= 1 two warps saw those barriers at the end of the previous
      // chunk's depthwise, in front of the block barrier that closed it -- no poll and no barrier of its own here
      if (c == 0) {
        if (warp == 0) mbar_wait(smem_u32(&bar_e), 0);
        if (warp == 3) {
          mbar_wait(smem_u32(&bar_w[0]), 0);
          if (n_chunks > 1) mbar_wait(smem_u32(&bar_w[1 % kWBuf]), 0);
        }
        __syncthreads();
      }
      par_e ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      MB_TICK(2);
      float em[16];
      int eb[16];                      // the NEXT chunk's bias (image c + 1 is here: warp 0 saw its barrier before the block barrier)
      load16(consts + 128 + half * 64, em);
      load16(s_wbuf + (uint32_t)((c + 1) % kWBuf) * a.img_stride + a.off_consts + half * 64, eb);
      {
        EeCtx e;
        e.tmem = tmem + lane_base + (uint32_t)(half * 16); e.dst = s_exp + (uint32_t)(half * 16) * cs;
        e.pos = &sPos[0][row]; e.n_win = a.n_win_tiles; e.wt0 = tgrp; e.wt_step = NT / 256;
        e.window_mask = window_mask; e.inside_mask = inside_mask; e.zpu = zpu;
        e.off = a.ex_off; e.span = a.ex_span; e.ulo = a.ex_ulo; e.fast = a.ex_rq.fast; e.store_next = c + 1 < n_chunks;
        switch (cs) {
          case 12: ee_tiles<12>(e, a.ex_rq, em, eb); break;
          case 20: ee_tiles<20>(e, a.ex_rq, em, eb); break;
          case 28: ee_tiles<28>(e, a.ex_rq, em, eb); break;
          case 36: ee_tiles<36>(e, a.ex_rq, em, eb); break;
          case 44: ee_tiles<44>(e, a.ex_rq, em, eb); break;
          case 52: ee_tiles<52>(e, a.ex_rq, em, eb); break;
          case 60: ee_tiles<60>(e, a.ex_rq, em, eb); break;
          case 68: ee_tiles<68>(e, a.ex_rq, em, eb); break;
          case 76: ee_tiles<76>(e, a.ex_rq, em, eb); break;
          case 84: ee_tiles<84>(e, a.ex_rq, em, eb); break;
          default: ee_tiles<0>(e, a.ex_rq, em, eb, cs);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n");
    }
    MB_TICK(3);
    __syncthreads();               // expanded planes complete (and free of the previous chunk's readers)
    MB_TICK(4);
    // ---- E(c + 1): runs on the tensor pipe while the SIMT pipes do this chunk's depthwise -----------------
    if (warp_u == 1 && a.has_expand && c + 1 < n_chunks) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      // image c + 1: warp 3 saw its barrier before this chunk's epilogue, two block barriers ago
      if (elect_one()) issue_expand(c + 1);
    }
    MB_TICK(5);
    // ---- DW + DE: lane = channel, warps walk strips of four output pixels -> middle planes (buffer c & 1) --
    {
      uint32_t wk[K][NWW];
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int j = 0; j < NWW; ++j) wk[ky][j] = ld_shared32(wb + a.off_taps + (uint32_t)((ky * NWW + j) * 32 + lane) * 4);
      const int dbias = (int)ld_shared32(consts + 256 + lane * 4);
      const float dmult = __uint_as_float(ld_shared32(consts + 384 + lane * 4));
      const uint32_t cbase = s_exp + (uint32_t)lane * cs;
      const uint32_t mbase = s_mid + (uint32_t)(c & 1) * mid_buf + (uint32_t)(lane >> 4) * a.mid_gstride + (uint32_t)(lane & 15);
      // strips are dealt starting at warp 2: when they do not divide evenly the extra ones miss warps 0 and 1,
      // which also issue the MMAs
      for (int sidx = (warp + NT / 32 - 2) % (NT / 32); sidx < a.n_strips; sidx += NT / 32) {
        const int ly = a.strips_x == 1 ? sidx : (int)__umulhi((uint32_t)sidx, a.inv_sx);
        const int x0 = (sidx - ly * a.strips_x) * 4;
        int acc[4] = {dbias, dbias, dbias, dbias};
        uint32_t raddr = cbase + (uint32_t)((ly * S) * a.row_stride + x0 * S);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const uint32_t w0 = ld_shared32(raddr), w1 = ld_shared32(raddr + 4);
          const uint32_t w2 = NLD == 3 ? ld_shared32(raddr + 8) : 0u;
          raddr += (uint32_t)a.row_stride;
          acc[0] = dp4a_us(window<0>(w0, w1, w2), wk[ky][0], acc[0]);
          acc[1] = dp4a_us(window<S>(w0, w1, w2), wk[ky][0], acc[1]);
          acc[2] = dp4a_us(window<2 * S>(w0, w1, w2), wk[ky][0], acc[2]);
          acc[3] = dp4a_us(window<3 * S>(w0, w1, w2), wk[ky][0], acc[3]);
          if (K == 5) {
            acc[0] = dp4a_us(byte_at<4>(w0, w1, w2), wk[ky][NWW - 1], acc[0]);
            acc[1] = dp4a_us(byte_at<S + 4>(w0, w1, w2), wk[ky][NWW - 1], acc[1]);
            acc[2] = dp4a_us(byte_at<2 * S + 4>(w0, w1, w2), wk[ky][NWW - 1], acc[2]);
            acc[3] = dp4a_us(byte_at<3 * S + 4>(w0, w1, w2), wk[ky][NWW - 1], acc[3]);
          }
        }
        const uint32_t y4 = a.dw_rq.fast ? a.dw_rq.pack4t<true>(acc[0], acc[1], acc[2], acc[3], dmult, dmult, dmult, dmult)
                                         : a.dw_rq.pack4t<false>(acc[0], acc[1], acc[2], acc[3], dmult, dmult, dmult, dmult);
        const uint32_t q = (uint32_t)(ly * a.TWp + x0);
        st_shared8(mbase + q * 16, y4);
        st_shared8(mbase + q * 16 + 16, y4 >> 8);
        st_shared8(mbase + q * 16 + 32, y4 >> 16);
        st_shared8(mbase + q * 16 + 48, y4 >> 24);
      }
    }
    MB_TICK(8);
    // the next chunk's barriers, polled by the two warps with the fewest strips (a completed try_wait costs ~100
    // cycles): E(c + 1) was issued at the start of this depthwise, image c + 2 a chunk ago
    if (a.has_expand && c + 1 < n_chunks) {
      if (warp == NT / 32 - 1) mbar_wait(smem_u32(&bar_e), par_e);
      if (warp == NT / 32 - 2 && c + 2 < n_chunks) mbar_wait(smem_u32(&bar_w[(c + 2) % kWBuf]), (uint32_t)(((c + 2) / kWBuf) & 1));
      // Buffer reuse: the middle buffer the NEXT depthwise writes and the weight buffer of image c + 3 were last read
      // by P(c - 1), issued a whole chunk ago: one poll frees both, and the image is requested three chunks ahead
      if (warp_u == NT / 32 - 3) {
        if (c >= 1) mbar_wait(smem_u32(&bar_p[(c - 1) & 1]), (uint32_t)(((c - 1) >> 1) & 1));
        if (c + 3 < n_chunks && elect_one()) load_image(c + 3);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    MB_TICK(9);
    // ---- P(c): project GEMM, accumulating over chunks ---------------------------------------------------
    if (warp_u == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      if (elect_one()) {
        const uint64_t bd = pj_bdesc + (uint64_t)(((uint32_t)(c % kWBuf) * a.img_stride) >> 4);
        for (int t = 0; t < a.n_out_tiles; ++t) {
          const uint64_t ad = pj_adesc + (uint64_t)((((uint32_t)(c & 1) * mid_buf) >> 4) + t * 128);
          const uint32_t d0 = tmem_u + (uint32_t)(a.col_pj + t * a.cout_p);
          umma_i8(d0, ad, bd, pj_id0, c > 0 ? 1u : 0u);
          if (pj_n0 < a.cout_p) umma_i8(d0 + (uint32_t)pj_n0, ad, bd + (uint64_t)((pj_n0 / 8) * 16), pj_id1, c > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_p[c & 1]));
      }
    }
    MB_TICK(10);
  }

  // ---- final epilogue: bias, requantise, residual, store ----------------------------------------------------
  if (warp == 0) mbar_wait(smem_u32(&bar_p[(n_chunks - 1) & 1]), (uint32_t)(((n_chunks - 1) >> 1) & 1));
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // columns this CTA finalises; with a partner, the partner's partial sums for them arrive in `stage`
  int col_lo = 0, col_hi = a.cout_p;
  const uint32_t s_stage = s_in + a.sm_stage;
  if (split == 2) {
    col_lo = rank == 0 ? 0 : a.col_split;
    col_hi = rank == 0 ? a.col_split : a.cout_p;
    const int p_lo = rank == 0 ? a.col_split : 0, p_hi = rank == 0 ? a.cout_p : a.col_split;
    cluster_sync_all();            // both CTAs are past their chunk loops: the staging areas (weight / plane buffers) are free
    const uint32_t remote = map_to_cta(s_stage, (uint32_t)(rank ^ 1));
    for (int t = tgrp; t < a.n_out_tiles; t += NT / 256)
      for (int c0 = p_lo + half * 16; c0 < p_hi; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + (uint32_t)(a.col_pj + t * a.cout_p + c0), v);
        const uint32_t dst = remote + (uint32_t)(t * 128 + row) * a.stage_stride + (uint32_t)(c0 - p_lo) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) st_cluster16(dst + j * 16, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    cluster_sync_all();            // partials delivered (release / acquire at cluster scope)
  }
  const int round = 1 << (a.add_shift > 0 ? a.add_shift - 1 : 0);
  for (int t = tgrp; t < a.n_out_tiles; t += NT / 256) {
    const int q = t * 128 + row;
    const int ly = (int)__umulhi((uint32_t)q, a.inv_twp);
    const int lx = q - ly * a.TWp;
    const int oy = oy0 + ly, ox = ox0 + lx;
    const bool valid = ly < a.TH && lx < a.TW && oy < a.Ho && ox < a.Wo;
    int8_t* o = a.out + (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.cout_p;
    // residual: the block input at the same pixel (stride 1), still in the input planes
    const uint32_t rpos = (uint32_t)((ly + a.pad_top) * a.WW + lx + a.pad_left) * 16;
    for (int c0 = col_lo + half * 16; c0 < col_hi; c0 += 32) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_base + (uint32_t)(a.col_pj + t * a.cout_p + c0), v);
      if (split == 2) {
        const uint32_t src = s_stage + (uint32_t)(t * 128 + row) * a.stage_stride + (uint32_t)(c0 - col_lo) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 pv = ld_shared16(src + j * 16);
          v[4 * j] += pv.x; v[4 * j + 1] += pv.y; v[4 * j + 2] += pv.z; v[4 * j + 3] += pv.w;
        }
      }
      if (!valid) continue;
      uint32_t packed[4];
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (a.has_res) rv = ld_shared16(s_in + (uint32_t)(c0 >> 4) * a.in_gstride + rpos);
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const int4 bq = *reinterpret_cast<const int4*>(sPjBias + c0 + w4 * 4);
        const float4 mq = *reinterpret_cast<const float4*>(sPjMult + c0 + w4 * 4);
        if (!a.has_res) {
          packed[w4] = a.pj_rq.pack4((int)v[w4 * 4 + 0] + bq.x, (int)v[w4 * 4 + 1] + bq.y, (int)v[w4 * 4 + 2] + bq.z,
                                     (int)v[w4 * 4 + 3] + bq.w, mq.x, mq.y, mq.z, mq.w);
        } else {
          const int bs[4] = {bq.x, bq.y, bq.z, bq.w};
          const float ms[4] = {mq.x, mq.y, mq.z, mq.w};
          int y[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[j] = a.pj_rq((int)v[w4 * 4 + j] + bs[j], ms[j]);
            const int r = (int)(int8_t)(rw[w4] >> (8 * j));
            const int s = (y[j] - a.pj_zp) * a.add_mult0 + (r - a.res_zp) * a.add_mult1 + round;
            y[j] = clampi((s >> a.add_shift) + a.zp_final, a.lo, a.hi);
          }
          packed[w4] = vbt::pack4_s8(y[0], y[1], y[2], y[3]);
        }
      }
      *reinterpret_cast<uint4*>(o + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
  MB_TICK(11);
  if (dbg_on)
    for (int i = 0; i < 12; ++i) a.dbg[(tid ? 12 : 0) + i] = dbg_acc[i];
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)a.tmem_cols));
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda): an int8 NHWC
// activation tensor as a 4-D byte tensor (cin_p, W, H, B), box = (16 bytes, WW, WH, 1): one 16-channel
// group of a tile's input window per load, 16-byte rows packed in (wy, wx) order.
bool encode_window_map(CUtensorMap* out, const void* base, int B, int H, int W, int cin_p, int WW, int WH) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeFn)p;
  }();
  if (!fn) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)cin_p, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)cin_p, (cuuint64_t)W * cin_p, (cuuint64_t)H * W * cin_p};
  const cuuint32_t box[4] = {16u, (cuuint32_t)WW, (cuuint32_t)WH, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

namespace vbt {

// ops: [expand PW (or null)] -> DW -> project PW (n_in == 2: + residual = the expand's input).
// *taken = false leaves the run to the single-op kernels.
int launch_mbconv_umma(const vbt_model* m, const OpRecord* ex, const OpRecord& dw, const OpRecord& pj,
                       const int8_t* in, int8_t* out, int B, cudaStream_t st, bool* taken) {
  static const bool enabled = [] { const char* e = getenv("VBT_MBCONV"); return !(e && e[0] == '0'); }();
  *taken = false;
  if (!enabled || dw.mb[0] <= 0) return VBT_OK;
  MbArgs a;
  a.in = in; a.out = out;
  a.img = m->dev_data + (size_t)(dw.mb[0] - 1) * 256;
  a.img_stride = dw.mb[1]; a.n_chunks = dw.mb[2];
  a.pj_bias = reinterpret_cast<const int32_t*>(m->dev_data + pj.bias_off);
  a.pj_mult = reinterpret_cast<const float*>(m->dev_data + pj.scale_off);
  a.B = B; a.H = dw.h_in; a.W = dw.w_in; a.Ho = dw.h_out; a.Wo = dw.w_out;
  a.has_expand = ex != nullptr;
  a.stem = ex && ex->type == OP_STEM;
  a.img_h = a.img_w = a.stem_pad_top = a.stem_pad_left = a.stem_zp = 0;
  if (a.stem) {
    a.img_h = ex->h_in; a.img_w = ex->w_in; a.stem_pad_top = ex->pad_top; a.stem_pad_left = ex->pad_left;
    a.stem_zp = ex->zp_in[0];
  }
  a.cin_p = a.stem ? 32 : (ex ? ex->cin_p : dw.cin_p);
  a.g_in = a.cin_p / 16; a.ge_in = (a.g_in + 1) / 2 * 2;
  a.cout_p = pj.cout_p;
  const int K = dw.k, S = dw.stride;
  a.pad_top = dw.pad_top; a.pad_left = dw.pad_left;
  a.has_res = pj.n_in == 2;
  if (a.cout_p > kMaxCout || a.cout_p % 16 || a.ge_in > 16 || (K != 3 && K != 5) || (S != 1 && S != 2)) return VBT_OK;
  if (!ex && (a.n_chunks != 1 || a.has_res)) return VBT_OK;
  if (a.has_res && (S != 1 || !ex || a.stem || pj.cout_p != ex->cin_p)) return VBT_OK;
  a.zp_fill = dw.zp_in[0];
  a.ex_off = a.ex_span = a.ex_ulo = 0;
  if (ex) {
    a.ex_rq = Requant(ex->zp_out, ex->act_lo, ex->act_hi, ex->requant_fast);
    a.ex_off = -(kRoundMagicBits + (ex->act_lo - ex->zp_out));
    a.ex_span = ex->act_hi - ex->act_lo;
    a.ex_ulo = ex->act_lo + 128;
  }
  a.dw_rq = Requant(dw.zp_out, dw.act_lo, dw.act_hi, dw.requant_fast);
  a.pj_rq = a.has_res ? Requant(pj.zp_out, -128, 127) : Requant(pj.zp_out, pj.act_lo, pj.act_hi, pj.requant_fast);
  a.pj_zp = pj.zp_out; a.res_zp = pj.zp_in[1];
  a.add_mult0 = pj.add_mult[0]; a.add_mult1 = pj.add_mult[1]; a.add_shift = pj.add_shift;
  a.zp_final = pj.zp_in[2]; a.lo = pj.act_lo; a.hi = pj.act_hi;
  // chunk image layout (effdet.mbconv_image_layout)
  const int sz_wexp = ex ? 32 * a.ge_in * 16 : 0;
  const int nww = K == 5 ? 2 : 1, nld = S == 1 ? 2 : 3;
  a.off_taps = sz_wexp;
  a.off_wproj = a.off_taps + K * nww * 128;          // [K rows][1 or 2 words][32 channels] u32
  a.off_consts = a.off_wproj + a.cout_p * 32;
  a.img_bytes = a.off_consts + 512;
  if (round_up(a.img_bytes, 128) != a.img_stride) {
    set_error("vbt_detect: MBConv weight image stride %d does not match the layout (%d)", a.img_stride, round_up(a.img_bytes, 128));
    return VBT_EFORMAT;
  }
  // ---- tile choice: TH x TW output pixels per CTA ------------------------------------------------------
  struct Geo { int TH, TW, TWp, WH, WW, m_total, n_win, n_out, strips_x, row_stride, chan_stride, cols; size_t smem; };
  auto geo = [&](int TH, int TW, Geo* g) {
    g->TH = TH; g->TW = TW; g->TWp = round_up(TW, 4);
    g->WH = (TH - 1) * S + K; g->WW = (TW - 1) * S + K;
    g->m_total = g->WH * g->WW;
    g->n_win = ex ? (g->m_total + 127) / 128 : 0;
    g->n_out = (TH * g->TWp + 127) / 128;
    g->strips_x = g->TWp / 4;
    // expanded tensor: [window row][channel][chan_stride bytes]; an odd number of words between channels, so that
    // lanes = channels of the depthwise hit 32 different banks
    g->chan_stride = round_up(std::max(g->WW, (g->strips_x - 1) * 4 * S + nld * 4), 4);
    if ((g->chan_stride / 4) % 2 == 0) g->chan_stride += 4;
    g->row_stride = 32 * g->chan_stride + 36;          // + 9 words: a warp of the epilogue that spans two window rows hits different banks
    g->cols = g->n_win * 32 + g->n_out * a.cout_p;
    g->smem = (size_t)a.ge_in * g->n_win * 2048 + (size_t)round_up(g->WH * g->row_stride, 128) +
              (size_t)4 * (g->n_out * 2048 + 16) + (size_t)kWBuf * a.img_stride + 128;
  };
  static const int max_cols = [] { const char* e = getenv("VBT_MB_MAXCOLS"); return e ? atoi(e) : 512; }();
  static const int env_tw = [] { const char* e = getenv("VBT_MB_TW"); return e ? atoi(e) : 0; }();
  static const int env_th = [] { const char* e = getenv("VBT_MB_TH"); return e ? atoi(e) : 0; }();
  long long best = -1;
  Geo bg = {};
  static std::map<std::tuple<const void*, int, int>, Geo> chosen;     // the search is per (op, batch), not per launch
  const auto key = std::make_tuple((const void*)&dw, B, (int)dw.mb[0]);
  const auto hit = chosen.find(key);
  const bool cached = hit != chosen.end();
  if (cached) { bg = hit->second; best = 0; }
  for (int TW = 4; !cached && TW <= a.Wo + 3; TW += 4) {
    const int tw = std::min(TW, a.Wo);
    if (env_tw && tw != std::min(env_tw, a.Wo)) continue;
    for (int TH = 1; TH <= a.Ho; ++TH) {
      if (env_th && TH != std::min(env_th, a.Ho)) continue;
      Geo g;
      geo(TH, tw, &g);
      if (g.n_out > 2 || g.n_win > 6 || g.cols > 512 || g.smem > 200 * 1024) continue;
      if (g.cols > max_cols && g.n_win * 32 + a.cout_p <= max_cols) continue;   // a narrower tile exists: keep TMEM for a co-resident CTA
      int cols_p = 32;
      while (cols_p < g.cols) cols_p <<= 1;
      // resident CTAs per SM: two by registers, three for tiles light enough for the 80-register build
      const int max_res = (cols_p <= 128 && g.smem <= 72 * 1024) ? 3 : 2;
      const int per_sm = std::max(1, std::min(std::min((int)(226 * 1024 / (g.smem + 4096)), 512 / cols_p), max_res));
      const long long ctas = (long long)B * ((a.Ho + TH - 1) / TH) * ((a.Wo + tw - 1) / tw);
      const long long waves = (ctas + 148LL * per_sm - 1) / (148LL * per_sm);
      // per-CTA cost model (~cycles): fill + per chunk (expand epilogue per window tile, planar depthwise per
      // strip, barriers) + final epilogue; two CTAs sharing an SM share its issue slots
      const long long strips = (long long)TH * g.strips_x;
      const long long per_chunk = 450 + 430LL * g.n_win + (strips + 7) / 8 * (K == 5 ? 140 : 70);
      // under-filled grids run as clusters of two CTAs that share a tile's chunks (below): half the chunk loop, plus the exchange
      const bool will_split = ex && a.n_chunks >= 4 && 2 * ctas <= 148LL * std::min(per_sm, 2);
      const long long my_chunks = will_split ? (a.n_chunks + 1) / 2 : a.n_chunks;
      const long long cta = 2500 + 40LL * g.n_win * a.ge_in + my_chunks * per_chunk + (long long)g.n_out * a.cout_p * (will_split ? 14 : 8);
      // CTAs that share an SM share its issue slots: two resident CTAs take ~1.7x one CTA's time
      const long long in_wave = std::min(ctas, 148LL * per_sm);
      const long long share = in_wave > 296 ? 22 : (in_wave > 148 ? 17 : 10);      // three resident CTAs: ~2.2x one CTA's time
      const long long cost = waves * cta * share / 10 * (cols_p > 256 ? 11 : 10) / 10;   // 512 columns: nothing else fits on the SM
      if (best < 0 || cost < best) { best = cost; bg = g; }
    }
  }
  if (best < 0) return VBT_OK;
  chosen[key] = bg;
  a.TH = bg.TH; a.TW = bg.TW; a.TWp = bg.TWp; a.WH = bg.WH; a.WW = bg.WW; a.m_total = bg.m_total;
  a.n_win_tiles = bg.n_win; a.n_out_tiles = bg.n_out; a.strips_x = bg.strips_x; a.n_strips = bg.TH * bg.strips_x;
  a.row_stride = bg.row_stride; a.chan_stride = bg.chan_stride;
  a.tiles_x = (a.Wo + a.TW - 1) / a.TW;
  const int tiles_y = (a.Ho + a.TH - 1) / a.TH;
  a.inv_ww = (uint32_t)((0x100000000ULL + a.WW - 1) / a.WW);
  a.inv_twp = (uint32_t)((0x100000000ULL + a.TWp - 1) / a.TWp);
  a.inv_gin = a.g_in > 1 ? (uint32_t)((0x100000000ULL + a.g_in - 1) / a.g_in) : 0u;
  a.inv_sx = a.strips_x > 1 ? (uint32_t)((0x100000000ULL + a.strips_x - 1) / a.strips_x) : 0u;
  a.in_gstride = (uint32_t)a.n_win_tiles * 2048;
  a.mid_gstride = (uint32_t)a.n_out_tiles * 2048 + 16;   // + 16: the two groups' planes land on different banks
  a.sm_exp = (uint32_t)a.ge_in * a.in_gstride * (ex ? 1 : 0);
  a.sm_mid = a.sm_exp + (uint32_t)round_up(a.WH * a.row_stride, 128);
  a.sm_wbuf = (uint32_t)round_up((int)(a.sm_mid + 4 * a.mid_gstride), 128);
  size_t smem = (size_t)a.sm_wbuf + (size_t)kWBuf * a.img_stride;
  a.col_pj = a.n_win_tiles * 32;
  int cols = 32;
  while (cols < a.col_pj + a.n_out_tiles * a.cout_p) cols <<= 1;
  a.tmem_cols = cols;
  if (cols > 512 || smem > 200 * 1024) return VBT_OK;
  // never more CTAs per SM than TMEM can serve, so tcgen05.alloc never spins
  smem = std::max(smem, (size_t)228 * 1024 / (512 / cols + 1));
  // Register budget: two CTAs per SM (111 registers) by default; tiles that need little shared memory and
  // TMEM (the single-chunk first block) run three per SM at 80 registers -- more warps to hide the
  // fill -> MMA -> epilogue latencies of a CTA that has only one chunk to pipeline
  static const bool occ3_on = [] { const char* e = getenv("VBT_MB_OCC3"); return !(e && e[0] == '0'); }();
  const bool occ3 = occ3_on && cols <= 128 && smem <= 72 * 1024;
  // one CTA per SM (512 TMEM columns or > 113 KB of shared memory): sixteen warps instead of eight hide the
  // latency chains of the epilogues and the depthwise that a second resident CTA would otherwise cover
  static const bool wide_on = [] { const char* e = getenv("VBT_MB_WIDE"); return e && e[0] == '1'; }();   // measured neutral: off
  const bool wide = wide_on && !occ3 && (cols > 256 || smem > 113 * 1024);
  void (*kern)(MbArgs, CUtensorMap);
  static const bool dbg = [] { const char* e = getenv("VBT_MB_DBG"); return e && e[0] == '1'; }();
#define MB_PICK(M, D) (K == 3 ? (S == 1 ? mbconv_umma_kernel<3, 1, M, D> : mbconv_umma_kernel<3, 2, M, D>) \
                              : (S == 1 ? mbconv_umma_kernel<5, 1, M, D> : mbconv_umma_kernel<5, 2, M, D>))
  if (dbg) kern = occ3 ? MB_PICK(3, true) : MB_PICK(2, true);          // the counters exist for the 256-thread builds only
  else kern = occ3 ? MB_PICK(3, false) : (wide ? MB_PICK(1, false) : MB_PICK(2, false));
  const int n_threads = (wide && !dbg) ? 512 : 256;
  static bool attr_set = false;
  if (!attr_set) {
    void (*all[20])(MbArgs, CUtensorMap) = {
        mbconv_umma_kernel<3, 1, 1, false>, mbconv_umma_kernel<3, 2, 1, false>, mbconv_umma_kernel<5, 1, 1, false>, mbconv_umma_kernel<5, 2, 1, false>,
        mbconv_umma_kernel<3, 1, 2, false>, mbconv_umma_kernel<3, 2, 2, false>, mbconv_umma_kernel<5, 1, 2, false>, mbconv_umma_kernel<5, 2, 2, false>,
        mbconv_umma_kernel<3, 1, 3, false>, mbconv_umma_kernel<3, 2, 3, false>, mbconv_umma_kernel<5, 1, 3, false>, mbconv_umma_kernel<5, 2, 3, false>,
        mbconv_umma_kernel<3, 1, 2, true>, mbconv_umma_kernel<3, 2, 2, true>, mbconv_umma_kernel<5, 1, 2, true>, mbconv_umma_kernel<5, 2, 2, true>,
        mbconv_umma_kernel<3, 1, 3, true>, mbconv_umma_kernel<3, 2, 3, true>, mbconv_umma_kernel<5, 1, 3, true>, mbconv_umma_kernel<5, 2, 3, true>};
    for (auto k : all) VBT_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  static long long* dbg_buf = nullptr;
  a.dbg = nullptr;
  if (dbg) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 24 * sizeof(long long));
    a.dbg = dbg_buf;
  }
  // Cluster split: when the grid leaves the GPU under-filled (one CTA per SM or fewer: the 20x20 and 10x10
  // stages at frame batch 64), two CTAs share a tile's chunks and exchange partial sums at the end.
  static const bool split_on = [] { const char* e = getenv("VBT_MB_SPLIT"); return !(e && e[0] == '0'); }();
  const long long n_ctas = (long long)a.tiles_x * tiles_y * B;
  a.split = 1; a.col_split = a.cout_p; a.stage_stride = 0; a.sm_stage = a.sm_exp;
  const int resident = std::max(1, std::min(std::min((int)(226 * 1024 / (smem + 4096)), 512 / cols), 2));   // CTAs per SM
  if (split_on && ex && a.n_chunks >= 4 && 2 * n_ctas <= 148LL * resident) {
    const int col_split = (a.cout_p / 32) * 16;
    const int stride = std::max(col_split, a.cout_p - col_split) * 4 + 16;     // + 16: rows land on different banks
    const size_t stage = (size_t)a.n_out_tiles * 128 * stride;
    const size_t need = (size_t)a.sm_exp + stage;
    if (std::max(need, smem) <= 200 * 1024) {
      a.split = 2; a.col_split = col_split; a.stage_stride = stride;
      smem = std::max(smem, need);
    }
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.tiles_x * tiles_y * a.split), (unsigned)B);
    cfg.blockDim = dim3(n_threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    static const bool pdl = [] { const char* e = getenv("VBT_MB_PDL"); return !(e && e[0] == '0'); }();
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (a.split == 2) {
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
      cfg.numAttrs = 2;
    }
    // the input window as tiled TMA loads: the activation tensor [B][H][W][cin_p] described to the copy engine
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    static const bool tma_on = [] { const char* e = getenv("VBT_MB_TMA"); return !(e && e[0] == '0'); }();
    a.use_tma = 0;
    if (tma_on && ex && !a.stem && a.WW <= 256 && a.WH <= 256) {
      static std::map<std::tuple<const void*, int, int, int, int, int, int>, CUtensorMap> maps;
      const auto mkey = std::make_tuple((const void*)in, B, a.H, a.W, a.cin_p, a.WW, a.WH);
      auto it = maps.find(mkey);
      if (it == maps.end()) {
        CUtensorMap tm;
        if (encode_window_map(&tm, in, B, a.H, a.W, a.cin_p, a.WW, a.WH)) it = maps.emplace(mkey, tm).first;
      }
      if (it != maps.end()) { tmap = it->second; a.use_tma = 1; }
    }
    VBT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a, tmap));
  }
  if (dbg) {       // debugging aid, never inside a graph capture: cycle counters of CTA (0, 0)
    long long h[24];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* nm[12] = {"prologue+fill", "wait image", "wait E", "EE", "sync", "E issue", "-", "-",
                                 "DW+DE", "fence+sync", "P issue", "final"};
    fprintf(stderr, "[mbconv %dx%d cin_p %d k%d s%d cout_p %d chunks %d | TH %d TW %d win_tiles %d out_tiles %d strips %d grid %d x %d split %d tma %d tmem %d smem %zu]\n",
            a.H, a.W, a.cin_p, K, S, a.cout_p, a.n_chunks, a.TH, a.TW, a.n_win_tiles, a.n_out_tiles, a.n_strips,
            a.tiles_x * tiles_y, B, a.split, a.use_tma, a.tmem_cols, smem);
    for (int i = 0; i < 12; ++i) fprintf(stderr, "   %-14s t0 %8lld   t64 %8lld\n", nm[i], h[i], h[12 + i]);
  }
  *taken = true;
  return VBT_OK;
}

}  // namespace vbt
