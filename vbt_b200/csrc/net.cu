// K2-K5 -- the EfficientDet-Lite network as int8 kernels over NHWC activations.
//
// replaces: the graph inside tflite_runtime's signature_fn(images=...) (odt.py:58-61):
// EfficientNet-Lite backbone (MBConv: 1x1 expand -> depthwise 3x3/5x5 -> 1x1 project
// [+ residual]), BiFPN (quantised sum fusion with nearest-up / max-pool-down resampling
// -> depthwise 3x3 -> 1x1) and the class / box heads, int8 per-tensor activations,
// per-channel weights, int32 accumulate, fp32 requantisation, fused ReLU6 clamps, the
// int8 LOGISTIC on the class output folded into the last conv's epilogue as a LUT.
//
// Data layout in HBM: every activation is [B, H, W, Cp] int8 with Cp = channels rounded
// up to 16 (pad channels hold the tensor's zero point and meet zero weights), so every
// pixel is a whole number of 128-bit words and a 1x1 conv is a K-major GEMM
// [B*H*W, Cin_p] x [Cout_p, Cin_p]^T.  Tensor t of the layer program lives at
// workspace + B * ws_offset[t].
//
// This file holds the SIMT kernels (stem, depthwise, fusion, pooling, and a dp4a
// pointwise kernel used for shapes the tcgen05 GEMM in pw_umma.cu does not take).
#include "model.cuh"
#include "requant.cuh"

namespace vbt {
cudaEvent_t* profile_begin(vbt_model* m);
int launch_pw_umma(const vbt_model* m, const OpRecord& op, const int8_t* in, const int8_t* res,
                   int8_t* out, long long out_batch_stride, int B, cudaStream_t st, bool* taken);
int launch_dw_umma(const vbt_model* m, const OpRecord& op, const int8_t* in, int8_t* out, int B,
                   cudaStream_t st, bool* taken);
int launch_node_umma(const vbt_model* m, const OpRecord* add0, const OpRecord* add, const OpRecord& dw,
                     const OpRecord& pw, const int8_t* const in[3], int8_t* out, long long out_batch_stride, int B,
                     cudaStream_t st, bool* taken);
int launch_mbconv_umma(const vbt_model* m, const OpRecord* ex, const OpRecord& dw, const OpRecord& pj,
                       const int8_t* in, int8_t* out, int B, cudaStream_t st, bool* taken);
}

namespace {

using vbt::OpRecord;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ int requant(int acc, float mult, int zp, int lo, int hi) {
  return clampi(__float2int_rn(__fmul_rn(__int2float_rn(acc), mult)) + zp, lo, hi);
}

__device__ __forceinline__ int s8(uint32_t word, int i) {
  return (int)(int8_t)(word >> (8 * i));
}

// ---------------------------------------------------------------------------------------
// stem: 3x3 stride-2 conv on the uint8 frame, 3 -> 32 channels, ReLU6 clamp.
// One thread = one output pixel x all output channels.  The 3x3 RGB window becomes nine
// (r,g,b,0) words; weights are stored the same way ([9][cout_p] words), so one
// dp4a.u32.s32 covers one pixel tap of one output channel.
// ---------------------------------------------------------------------------------------
struct StemArgs {
  const uint8_t* in; int8_t* out;
  const uint32_t* w; const int32_t* bias; const float* mult;   // w [9][cout_p]
  int B, H, W, Ho, Wo, cout_p, pad_top, pad_left, zp_in;
  vbt::Requant rq;
};

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__global__ void __launch_bounds__(128) stem_kernel(StemArgs a) {
  __shared__ __align__(16) uint32_t sw[9 * 64];
  __shared__ __align__(16) int32_t sb[64];
  __shared__ __align__(16) float sm[64];
  for (int i = threadIdx.x; i < a.cout_p * 9; i += blockDim.x) sw[i] = a.w[i];
  for (int i = threadIdx.x; i < a.cout_p; i += blockDim.x) { sb[i] = a.bias[i]; sm[i] = a.mult[i]; }
  __syncthreads();
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();
  const long long total = (long long)a.B * a.Ho * a.Wo;
  const uint32_t zpix = (uint32_t)a.zp_in * 0x010101u;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % a.Wo), oy = (int)((p / a.Wo) % a.Ho), b = (int)(p / ((long long)a.Wo * a.Ho));
    uint32_t x[9];
    const uint8_t* fin = a.in + (size_t)b * a.H * a.W * 3;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - a.pad_top + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - a.pad_left + kx;
        uint32_t v = zpix;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
          const uint8_t* px = fin + ((size_t)iy * a.W + ix) * 3;
          v = (uint32_t)__ldg(px) | ((uint32_t)__ldg(px + 1) << 8) | ((uint32_t)__ldg(px + 2) << 16);
        }
        x[ky * 3 + kx] = v;
      }
    }
    int8_t* o = a.out + (size_t)p * a.cout_p;
    for (int c0 = 0; c0 < a.cout_p; c0 += 16) {
      int acc[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 bv = *reinterpret_cast<const int4*>(sb + c0 + q * 4);
        acc[q * 4 + 0] = bv.x; acc[q * 4 + 1] = bv.y; acc[q * 4 + 2] = bv.z; acc[q * 4 + 3] = bv.w;
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 wv = *reinterpret_cast<const uint4*>(sw + t * a.cout_p + c0 + q * 4);
          acc[q * 4 + 0] = dp4a_us(x[t], wv.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = dp4a_us(x[t], wv.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = dp4a_us(x[t], wv.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = dp4a_us(x[t], wv.w, acc[q * 4 + 3]);
        }
      }
      uint32_t packed[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 mv = *reinterpret_cast<const float4*>(sm + c0 + q * 4);
        packed[q] = a.rq.pack4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3], mv.x, mv.y, mv.z, mv.w);
      }
      *reinterpret_cast<uint4*>(o + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// depthwise KxK (K = 3 or 5, stride 1 or 2, TF-SAME), channel-word stationary.
//
// One thread owns ONE 32-bit activation word (4 channels) and keeps everything that
// depends only on the channel in registers for its whole life: the K*K*4 pre-masked weight
// words (channel j of the word in byte j, see effdet.pack_blob -- dp4a(x_word, w_word)
// multiplies exactly one channel, no unpacking), 4 biases, 4 multipliers.  It then walks
// row segments of L output pixels in groups of J: the (J-1)*S+K window columns of a group
// are loaded together (32-bit loads), then each output is K*K*4 dp4a.  Consecutive
// threads hold consecutive channel words, so every load / store of a warp is one contiguous
// 128-byte line (or whole 32-byte sectors when the tensor has fewer than 128 channels):
// the L1 wavefront count per output drops ~10x against a 16-channel-per-thread layout.
// ---------------------------------------------------------------------------------------
struct DwArgs {
  const int8_t* in; int8_t* out;
  const uint32_t* w; const int32_t* bias; const float* mult;   // w [k*k][c_p] masked words
  int B, H, W, Ho, Wo, c_p, pad_top, pad_left, zp_in;
  int words, wx, seg_len, n_seg;       // c_p/4, word lanes per CTA, outputs per segment, segments per row
  long long n_items;                   // B * Ho * n_seg
  vbt::Requant rq;
};

template <int K, int S, bool FAST>
__global__ void __launch_bounds__(256) dw_kernel(DwArgs a) {
  constexpr int J = (K == 5) ? (S == 1 ? 4 : 2) : (S == 1 ? 4 : 2);     // outputs per group
  constexpr bool kSmemW = (K == 5);   // 5x5: the 100 weight words live in shared memory, not registers
  constexpr int NCOL = (J - 1) * S + K;
  const int wl = threadIdx.x % a.wx, il = threadIdx.x / a.wx;
  const int il_n = blockDim.x / a.wx;
  const int cw = blockIdx.y * a.wx + wl;
  if (cw >= a.words || il >= il_n) return;
  // channel-stationary state
  extern __shared__ __align__(16) uint4 sw[];         // [K*K][blockDim.x] when kSmemW
  uint32_t w[kSmemW ? 1 : K * K][4];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w + (size_t)t * a.c_p) + cw);
    if (kSmemW) {
      sw[t * blockDim.x + threadIdx.x] = v;           // read back only by this thread: no barrier
    } else {
      w[t][0] = v.x; w[t][1] = v.y; w[t][2] = v.z; w[t][3] = v.w;
    }
  }
  const int4 bias = __ldg(reinterpret_cast<const int4*>(a.bias) + cw);
  const float4 mult = __ldg(reinterpret_cast<const float4*>(a.mult) + cw);
  const uint32_t zpw = (uint32_t)(a.zp_in & 0xff) * 0x01010101u;
  // All addressing is a warp-uniform base pointer + a 32-bit word offset (tensors are far
  // below 2^32 words), so a load costs one integer add and the LDG itself.
  const uint32_t* __restrict__ in = reinterpret_cast<const uint32_t*>(a.in);
  uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(a.out);
  const uint32_t words = (uint32_t)a.words;
  const uint32_t row_words = (uint32_t)a.W * words;
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();

  const int n_items = (int)a.n_items;                 // < 2^31: 32-bit index arithmetic
  for (int item = blockIdx.x * il_n + il; item < n_items; item += gridDim.x * il_n) {
    const int seg = item % a.n_seg;
    const int r = item / a.n_seg;
    const int oy = r % a.Ho, b = r / a.Ho;
    const int x0 = seg * a.seg_len, x1 = min(x0 + a.seg_len, a.Wo);
    const uint32_t frame = (uint32_t)b * (uint32_t)a.H * row_words + (uint32_t)cw;
    uint32_t row_off[K];
    bool row_ok[K];
    bool rows_in = true;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * S - a.pad_top + ky;
      row_ok[ky] = iy >= 0 && iy < a.H;
      rows_in = rows_in && row_ok[ky];
      row_off[ky] = frame + (uint32_t)(row_ok[ky] ? iy : 0) * row_words;
    }
    // J outputs per group: all (J-1)*S+K window columns are loaded first (independent loads in
    // flight together), then the J outputs are computed from registers
    uint32_t out_off = (((uint32_t)b * a.Ho + oy) * a.Wo + x0) * words + (uint32_t)cw;
    for (int xg = x0; xg < x1; xg += J) {
      uint32_t win[NCOL][K];
      const int ixb = xg * S - a.pad_left;
      if (rows_in && ixb >= 0 && ixb + NCOL <= a.W) {       // interior: no predicates
        uint32_t col = (uint32_t)ixb * words;
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
#pragma unroll
          for (int ky = 0; ky < K; ++ky) win[c][ky] = __ldg(in + (row_off[ky] + col));
          col += words;
        }
      } else {                                              // image border: zero-point padding
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
          const int ix = ixb + c;
          const bool col_ok = ix >= 0 && ix < a.W;
          const uint32_t col = (uint32_t)(col_ok ? ix : 0) * words;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
            const uint32_t v = __ldg(in + (row_off[ky] + col));
            win[c][ky] = (col_ok && row_ok[ky]) ? v : zpw;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int x = xg + j;
        if (x < x1) {
          int acc0 = bias.x, acc1 = bias.y, acc2 = bias.z, acc3 = bias.w;
#pragma unroll
          for (int ky = 0; ky < K; ++ky) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
              const int xv = (int)win[j * S + kx][ky];
              const int t = ky * K + kx;
              uint4 wv;
              if (kSmemW) wv = sw[t * blockDim.x + threadIdx.x];
              else wv = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
              acc0 = __dp4a(xv, (int)wv.x, acc0);
              acc1 = __dp4a(xv, (int)wv.y, acc1);
              acc2 = __dp4a(xv, (int)wv.z, acc2);
              acc3 = __dp4a(xv, (int)wv.w, acc3);
            }
          }
          out[out_off] = a.rq.pack4t<FAST>(acc0, acc1, acc2, acc3, mult.x, mult.y, mult.z, mult.w);
        }
        out_off += words;
      }
    }
  }
}

template <int K, int S>
int launch_dw(DwArgs a, cudaStream_t st) {
  constexpr int kMaxThreads = (K == 5) ? 128 : 256;   // 5x5: ~145 registers per thread
  a.words = a.c_p / 4;
  const int chunks = (a.words + kMaxThreads - 1) / kMaxThreads;
  a.wx = (a.words + chunks - 1) / chunks;
  int il_n = kMaxThreads / a.wx;
  if (il_n < 1) il_n = 1;
  const int threads = a.wx * il_n;
  // segment length: long enough to amortise the K-S warm-up columns, short enough to give
  // every SM several CTAs of work
  a.seg_len = a.Wo <= 24 ? a.Wo : 16;
  a.n_seg = (a.Wo + a.seg_len - 1) / a.seg_len;
  a.n_items = (long long)a.B * a.Ho * a.n_seg;
  long long gx = (a.n_items + il_n - 1) / il_n;
  const long long cap = 148LL * 8;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)chunks);
  const size_t smem = (K == 5) ? (size_t)K * K * threads * sizeof(uint4) : 0;
  if (smem > 48 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      VBT_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel<K, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      VBT_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel<K, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      attr_set = true;
    }
  }
  if (a.rq.fast) VBT_CHECK_CUDA(vbt::launch_pdl(dw_kernel<K, S, true>, grid, dim3(threads), smem, st, a));
  else VBT_CHECK_CUDA(vbt::launch_pdl(dw_kernel<K, S, false>, grid, dim3(threads), smem, st, a));
  return VBT_OK;
}

// ---------------------------------------------------------------------------------------
// BiFPN fusion: quantised sum of up to 3 inputs, each optionally resampled to the node's
// level (nearest-neighbour up, 3x3 stride-2 SAME max-pool down), ReLU6 clamp.
// Also serves the stand-alone max-pool (n_in == 1, identity rescale skipped).
// ---------------------------------------------------------------------------------------
struct AddArgs {
  const int8_t* in[3]; int8_t* out;
  int n_in, B, Ho, Wo, c_p;
  int in_h[3], in_w[3], resample[3], zp_in[3], mult[3];
  int shift, zp_out, lo, hi, pool_only;
};

__device__ __forceinline__ uint4 fetch_resampled(const int8_t* base, int b, int oy, int ox, int Ho,
                                                 int Wo, int ih, int iw, int mode, int c_p, int c0) {
  const int8_t* fin = base + (size_t)b * ih * iw * c_p + c0;
  if (mode == vbt::RS_NONE) {
    return __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)oy * iw + ox) * c_p));
  } else if (mode == vbt::RS_UP_NEAREST) {
    const int sy = (oy * ih) / Ho, sx = (ox * iw) / Wo;
    return __ldg(reinterpret_cast<const uint4*>(fin + ((size_t)sy * iw + sx) * c_p));
  }
  // 3x3 stride-2 SAME max-pool: pad_before = total/2, out-of-range taps are skipped
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

__global__ void __launch_bounds__(256) add_kernel(AddArgs a) {
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();
  const int groups = a.c_p >> 4;
  const long long total = (long long)a.B * a.Ho * a.Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(i % groups);
    const long long p = i / groups;
    const int ox = (int)(p % a.Wo), oy = (int)((p / a.Wo) % a.Ho), b = (int)(p / ((long long)a.Wo * a.Ho));
    const int c0 = gidx << 4;
    uint4 o;
    if (a.pool_only) {
      o = fetch_resampled(a.in[0], b, oy, ox, a.Ho, a.Wo, a.in_h[0], a.in_w[0], vbt::RS_DOWN_MAXPOOL,
                          a.c_p, c0);
    } else {
      int acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 1 << (a.shift - 1);
      for (int n = 0; n < a.n_in; ++n) {
        const uint4 v = fetch_resampled(a.in[n], b, oy, ox, a.Ho, a.Wo, a.in_h[n], a.in_w[n],
                                        a.resample[n], a.c_p, c0);
        const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[q * 4 + j] += (s8(xs[q], j) - a.zp_in[n]) * a.mult[n];
      }
      uint32_t packed[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int y = clampi((acc[j] >> a.shift) + a.zp_out, a.lo, a.hi);
        packed[j >> 2] |= (uint32_t)(y & 0xff) << (8 * (j & 3));
      }
      o = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    *reinterpret_cast<uint4*>(a.out + (size_t)p * a.c_p + c0) = o;
  }
}

// ---------------------------------------------------------------------------------------
// pointwise 1x1 conv, SIMT dp4a GEMM: 64 pixels x 64 output channels per CTA, K chunks
// of 64 bytes staged in shared memory; epilogue = requantise (+ quantised residual add)
// (+ LOGISTIC LUT), written through shared memory as 128-bit rows when the output is a
// padded workspace tensor, bytewise for the packed head outputs.
// ---------------------------------------------------------------------------------------
struct PwArgs {
  const int8_t* in; const int8_t* res; int8_t* out;
  const int8_t* w; const int32_t* bias; const float* mult; const int8_t* lut;
  long long M;                // B * H * W
  int HW, cin_p, cout, cout_p;
  int zp_conv, lo, hi;        // requant target of the conv itself
  int has_res, res_zp, add_mult0, add_mult1, add_shift, zp_final;
  int out_pix_stride; long long out_batch_stride, out_elem_offset;
  int vector_out;             // 1: out rows are 16-byte aligned multiples
};

constexpr int PW_BM = 64, PW_BN = 64, PW_BK = 64;

__global__ void __launch_bounds__(256) pw_dp4a_kernel(PwArgs a) {
  __shared__ int As[PW_BM][PW_BK / 4 + 1];
  __shared__ int Ws[PW_BN][PW_BK / 4 + 1];
  __shared__ __align__(16) int8_t Os[PW_BM][PW_BN];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * PW_BM;
  const int n0 = blockIdx.y * PW_BN;
  const int ty = tid >> 4, tx = tid & 15;     // rows ty*4..+3 ; cols tx + 16*j
  vbt::pdl_wait();
  vbt::pdl_launch_dependents();
  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  const int lrow = tid >> 2, lseg = tid & 3;  // loader: one 16-byte segment per thread
  for (int k0 = 0; k0 < a.cin_p; k0 += PW_BK) {
    int4 av = make_int4(0, 0, 0, 0), wv = make_int4(0, 0, 0, 0);
    const int kk = k0 + lseg * 16;
    if (kk < a.cin_p) {
      if (m0 + lrow < a.M)
        av = __ldg(reinterpret_cast<const int4*>(a.in + (size_t)(m0 + lrow) * a.cin_p + kk));
      if (n0 + lrow < a.cout_p)
        wv = __ldg(reinterpret_cast<const int4*>(a.w + (size_t)(n0 + lrow) * a.cin_p + kk));
    }
    __syncthreads();
    As[lrow][lseg * 4 + 0] = av.x; As[lrow][lseg * 4 + 1] = av.y;
    As[lrow][lseg * 4 + 2] = av.z; As[lrow][lseg * 4 + 3] = av.w;
    Ws[lrow][lseg * 4 + 0] = wv.x; Ws[lrow][lseg * 4 + 1] = wv.y;
    Ws[lrow][lseg * 4 + 2] = wv.z; Ws[lrow][lseg * 4 + 3] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kw = 0; kw < PW_BK / 4; ++kw) {
      int av4[4], wv4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av4[i] = As[ty * 4 + i][kw];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv4[j] = Ws[tx + 16 * j][kw];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a(av4[i], wv4[j], acc[i][j]);
    }
  }
  // epilogue
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx + 16 * j;
    const bool nvalid = n < a.cout_p;
    const int bias = nvalid ? __ldg(a.bias + n) : 0;
    const float mult = nvalid ? __ldg(a.mult + n) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ty * 4 + i;
      int y = requant(acc[i][j] + bias, mult, a.zp_conv, a.has_res ? -128 : a.lo, a.has_res ? 127 : a.hi);
      if (a.has_res && nvalid && m < a.M) {
        const int r = (int)a.res[(size_t)m * a.cout_p + n];
        const int s = (y - a.zp_conv) * a.add_mult0 + (r - a.res_zp) * a.add_mult1 +
                      (1 << (a.add_shift - 1));
        y = clampi((s >> a.add_shift) + a.zp_final, a.lo, a.hi);
      }
      if (a.lut) y = (int)a.lut[y + 128];
      Os[ty * 4 + i][tx + 16 * j] = (int8_t)y;
    }
  }
  __syncthreads();
  if (a.vector_out) {
    const long long m = m0 + lrow;
    const int n = n0 + lseg * 16;
    if (m < a.M && n < a.cout_p) {
      const long long b = m / a.HW, p = m % a.HW;
      int8_t* dst = a.out + b * a.out_batch_stride + a.out_elem_offset + p * a.out_pix_stride + n;
      *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(&Os[lrow][lseg * 16]);
    }
  } else {
    for (int e = tid; e < PW_BM * PW_BN; e += 256) {
      const int r = e / PW_BN, c = e % PW_BN;
      const long long m = m0 + r;
      const int n = n0 + c;
      if (m < a.M && n < a.cout) {
        const long long b = m / a.HW, p = m % a.HW;
        a.out[b * a.out_batch_stride + a.out_elem_offset + p * a.out_pix_stride + n] = Os[r][c];
      }
    }
  }
}

int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

namespace {

// Enqueue the whole layer program on `st` (one launch per op).  `prof`: event set for the
// per-op timing mode (bench.py), else nullptr.
int run_ops(vbt_model* m, const uint8_t* dev_in, int B, void* dev_workspace, int8_t* dev_out_cls,
            int8_t* dev_out_box, cudaStream_t main_st, cudaEvent_t* prof) {
  using namespace vbt;
  uint8_t* ws = static_cast<uint8_t*>(dev_workspace);
  const long long Np = m->hdr.n_anchors_pad;
  // tensors without a workspace slot (the model input) live in the caller's buffer
  auto tensor_ptr = [&](int id) -> int8_t* {
    const int64_t off = m->tensors[id].ws_offset;
    if (off < 0) return reinterpret_cast<int8_t*>(const_cast<uint8_t*>(dev_in));
    return reinterpret_cast<int8_t*>(ws + (size_t)B * (size_t)off);
  };
  auto data = [&](int64_t off) { return off < 0 ? nullptr : m->dev_data + off; };
  int launched = 0;
  m->kernels_last_run = 0;
  // Branch k > 0 (a head chain) runs on its own stream: forked once the trunk is enqueued,
  // joined at the end.  Under stream capture this becomes parallel branches of the graph.
  const bool fork = !prof && m->n_branches > 0;
  bool started[vbt_model::kMaxBranches] = {};
  if (prof) VBT_CHECK_CUDA(cudaEventRecord(prof[0], main_st));
  const int n_ops = (int)m->ops.size();
  // scheduler counters of this workspace (a lane): -1 = all slots taken, persistent kernels stay off
  int cslot = -1;
  for (size_t i = 0; i < m->counter_owner.size(); ++i)
    if (m->counter_owner[i] == dev_workspace) cslot = (int)i;
  if (cslot < 0 && (int)m->counter_owner.size() < vbt_model::kCounterSlots) {
    m->counter_owner.push_back(dev_workspace);
    cslot = (int)m->counter_owner.size() - 1;
  }
  for (int oi = 0; oi < n_ops; ++oi) {
    const OpRecord& op = m->ops[oi];
    m->cur_counters = (cslot >= 0 && m->dev_counters) ? m->dev_counters + ((size_t)cslot * n_ops + oi) * 16 : nullptr;
    cudaStream_t st = main_st;
    if (fork && op.branch > 0) {
      const int k = op.branch - 1;
      if (!started[k]) { VBT_CHECK_CUDA(cudaStreamWaitEvent(m->branch_stream[k], m->fork_event[k], 0)); started[k] = true; }
      st = m->branch_stream[k];
    }
    // fused [ADD ->] DW3x3 -> PW group: one kernel; falls through to the single ops if declined
    int covered = 1;
    if (m->fuse[oi] > 1 && m->fuse_kind[oi] == 1) {
      // MBConv block: [expand ->] depthwise -> project as one kernel (csrc/mbconv_umma.cu)
      const int glen = m->fuse[oi];
      const OpRecord* ex = glen == 3 ? &op : nullptr;
      const OpRecord& dwop = m->ops[oi + glen - 2];
      const OpRecord& pjop = m->ops[oi + glen - 1];
      bool mb_taken = false;
      if (int rc = launch_mbconv_umma(m, ex, dwop, pjop, tensor_ptr(op.in[0]), tensor_ptr(pjop.out), B, st, &mb_taken)) return rc;
      if (mb_taken) covered = glen;
    } else if (m->fuse[oi] > 1) {
      const int glen = m->fuse[oi];                        // 2: DW PW; 3: ADD DW PW; 4: ADD ADD DW PW
      const OpRecord* add0 = glen == 4 ? &op : nullptr;
      const OpRecord* add = glen >= 3 ? &m->ops[oi + glen - 3] : nullptr;
      const OpRecord& dwop = m->ops[oi + glen - 2];
      const OpRecord& pwop = m->ops[oi + glen - 1];
      const int8_t* ins[3] = {nullptr, nullptr, nullptr};
      if (add0) {
        ins[0] = tensor_ptr(add0->in[0]); ins[1] = tensor_ptr(add0->in[1]);
        ins[2] = tensor_ptr(add->in[0] == add0->out ? add->in[1] : add->in[0]);
      } else if (add) {
        for (int i = 0; i < add->n_in; ++i) ins[i] = tensor_ptr(add->in[i]);
      } else {
        ins[0] = tensor_ptr(dwop.in[0]);
      }
      int8_t* out;
      long long obs;
      if (pwop.out_kind == 0) { out = tensor_ptr(pwop.out); obs = (long long)pwop.h_out * pwop.w_out * pwop.cout_p; }
      else if (pwop.out_kind == 1) { out = dev_out_cls; obs = Np * m->hdr.n_classes; }
      else { out = dev_out_box; obs = Np * 4; }
      bool node_taken = false;
      if (int rc = launch_node_umma(m, add0, add, dwop, pwop, ins, out, obs, B, st, &node_taken)) return rc;
      if (node_taken) covered = glen;
    }
    if (covered == 1) {
    switch (op.type) {
      case OP_STEM: {
        StemArgs a;
        a.in = dev_in; a.out = tensor_ptr(op.out);
        a.w = reinterpret_cast<const uint32_t*>(data(op.w_off));
        a.bias = reinterpret_cast<const int32_t*>(data(op.bias_off));
        a.mult = reinterpret_cast<const float*>(data(op.scale_off));
        a.B = B; a.H = op.h_in; a.W = op.w_in; a.Ho = op.h_out; a.Wo = op.w_out;
        a.cout_p = op.cout_p; a.pad_top = op.pad_top; a.pad_left = op.pad_left;
        a.zp_in = op.zp_in[0]; a.rq = Requant(op.zp_out, op.act_lo, op.act_hi, op.requant_fast);
        VBT_REQUIRE(op.cout_p <= 64 && op.k == 3 && op.stride == 2, "vbt_detect: unsupported stem");
        VBT_CHECK_CUDA(launch_pdl(stem_kernel, dim3(grid_for((long long)B * op.h_out * op.w_out, 128)), dim3(128), 0, st, a));
        break;
      }
      case OP_DW: {
        bool dw_taken = false;
        if (int rc = launch_dw_umma(m, op, tensor_ptr(op.in[0]), tensor_ptr(op.out), B, st, &dw_taken)) return rc;
        if (dw_taken) break;
        DwArgs a;
        a.in = tensor_ptr(op.in[0]); a.out = tensor_ptr(op.out);
        a.w = reinterpret_cast<const uint32_t*>(data(op.w_off));
        a.bias = reinterpret_cast<const int32_t*>(data(op.bias_off));
        a.mult = reinterpret_cast<const float*>(data(op.scale_off));
        a.B = B; a.H = op.h_in; a.W = op.w_in; a.Ho = op.h_out; a.Wo = op.w_out; a.c_p = op.cout_p;
        a.pad_top = op.pad_top; a.pad_left = op.pad_left;
        a.zp_in = op.zp_in[0]; a.rq = Requant(op.zp_out, op.act_lo, op.act_hi, op.requant_fast);
        int rc = VBT_OK;
        if (op.k == 3 && op.stride == 1) rc = launch_dw<3, 1>(a, st);
        else if (op.k == 3 && op.stride == 2) rc = launch_dw<3, 2>(a, st);
        else if (op.k == 5 && op.stride == 1) rc = launch_dw<5, 1>(a, st);
        else if (op.k == 5 && op.stride == 2) rc = launch_dw<5, 2>(a, st);
        else VBT_REQUIRE(false, "vbt_detect: depthwise kernel %dx%d stride %d", op.k, op.k, op.stride);
        if (rc) return rc;
        break;
      }
      case OP_ADD:
      case OP_MAXPOOL: {
        AddArgs a;
        a.n_in = op.n_in; a.B = B; a.Ho = op.h_out; a.Wo = op.w_out; a.c_p = op.cout_p;
        for (int i = 0; i < 3; ++i) {
          a.in[i] = (i < op.n_in) ? tensor_ptr(op.in[i]) : nullptr;
          a.in_h[i] = op.in_h[i]; a.in_w[i] = op.in_w[i]; a.resample[i] = op.resample[i];
          a.zp_in[i] = op.zp_in[i]; a.mult[i] = op.add_mult[i];
        }
        a.shift = op.add_shift; a.zp_out = op.zp_out; a.lo = op.act_lo; a.hi = op.act_hi;
        a.pool_only = (op.type == OP_MAXPOOL);
        a.out = tensor_ptr(op.out);
        VBT_CHECK_CUDA(launch_pdl(add_kernel, dim3(grid_for((long long)B * op.h_out * op.w_out * (op.cout_p / 16), 256)), dim3(256), 0, st, a));
        break;
      }
      case OP_PW: {
        VBT_REQUIRE(op.pw_dtype == 0, "vbt_detect: bf16 pointwise op %d is not inside a fused head stage "
                    "(csrc/node_umma.cu); there is no stand-alone bf16 kernel", oi);
        const int8_t* in = tensor_ptr(op.in[0]);
        const int8_t* res = (op.n_in == 2) ? tensor_ptr(op.in[1]) : nullptr;
        int8_t* out;
        long long out_batch_stride;
        if (op.out_kind == 0) { out = tensor_ptr(op.out); out_batch_stride = (long long)op.h_out * op.w_out * op.cout_p; }
        else if (op.out_kind == 1) { out = dev_out_cls; out_batch_stride = Np * m->hdr.n_classes; }
        else { out = dev_out_box; out_batch_stride = Np * 4; }
        bool taken = false;
        if (int rc = launch_pw_umma(m, op, in, res, out, out_batch_stride, B, st, &taken)) return rc;
        if (!taken) {
          PwArgs a;
          a.in = in; a.res = res; a.out = out;
          a.w = reinterpret_cast<const int8_t*>(data(op.w_off));
          a.bias = reinterpret_cast<const int32_t*>(data(op.bias_off));
          a.mult = reinterpret_cast<const float*>(data(op.scale_off));
          a.lut = reinterpret_cast<const int8_t*>(data(op.lut_off));
          a.HW = op.h_in * op.w_in; a.M = (long long)B * a.HW;
          a.cin_p = op.cin_p; a.cout = op.cout; a.cout_p = op.cout_p;
          a.zp_conv = op.zp_out; a.lo = op.act_lo; a.hi = op.act_hi;
          a.has_res = res != nullptr; a.res_zp = op.zp_in[1];
          a.add_mult0 = op.add_mult[0]; a.add_mult1 = op.add_mult[1]; a.add_shift = op.add_shift;
          a.zp_final = op.zp_in[2];
          a.out_pix_stride = op.out_pix_stride; a.out_batch_stride = out_batch_stride;
          a.out_elem_offset = op.out_elem_offset;
          a.vector_out = (op.out_kind == 0);
          dim3 grid((unsigned)((a.M + PW_BM - 1) / PW_BM), (unsigned)((op.cout_p + PW_BN - 1) / PW_BN));
          VBT_CHECK_CUDA(launch_pdl(pw_dp4a_kernel, grid, dim3(256), 0, st, a));
        }
        break;
      }
      default:
        VBT_REQUIRE(false, "vbt_detect: unknown op type %d", op.type);
    }
    }
    VBT_CHECK_CUDA(cudaPeekAtLastError());
    ++m->kernels_last_run;
    for (int c = 0; c < covered; ++c) {              // a group's time lands on its first op
      ++launched;
      if (prof) VBT_CHECK_CUDA(cudaEventRecord(prof[launched], main_st));
      if (fork && op.branch == 0)
        for (int k = 0; k < m->n_branches; ++k)
          if (m->fork_after[k] == launched - 1) VBT_CHECK_CUDA(cudaEventRecord(m->fork_event[k], main_st));
    }
    oi += covered - 1;
  }
  for (int k = 0; k < vbt_model::kMaxBranches; ++k) {
    if (!started[k]) continue;
    VBT_CHECK_CUDA(cudaEventRecord(m->join_event[k], m->branch_stream[k]));
    VBT_CHECK_CUDA(cudaStreamWaitEvent(main_st, m->join_event[k], 0));
  }
  return VBT_OK;
}

}  // namespace

extern "C" int vbt_detect(vbt_model* m, const uint8_t* dev_in, int B, void* dev_workspace,
                          size_t workspace_bytes, int8_t* dev_out_cls, int8_t* dev_out_box,
                          void* stream) {
  using namespace vbt;
  VBT_REQUIRE(m && dev_in && dev_workspace && dev_out_cls && dev_out_box, "vbt_detect: null pointer");
  VBT_REQUIRE(B > 0, "vbt_detect: B=%d", B);
  VBT_REQUIRE(workspace_bytes >= (size_t)B * (size_t)m->hdr.ws_bytes_per_frame,
              "vbt_detect: workspace of %zu bytes is smaller than B * %lld", workspace_bytes,
              (long long)m->hdr.ws_bytes_per_frame);
  VBT_REQUIRE(((uintptr_t)dev_workspace & 255) == 0, "vbt_detect: workspace must be 256-byte aligned");
  VBT_REQUIRE(!m->ops.empty(), "vbt_detect: the model holds no layer program (anchors only)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_ops = (int)m->ops.size();
  cudaEvent_t* prof = profile_begin(m);
  // Replaying a captured CUDA graph turns the ~200 launches of a batch into one submission.
  // Not on the legacy default stream (capture is illegal there), not while per-op timing is on.
  static const bool use_graph = [] { const char* e = getenv("VBT_GRAPH"); return !(e && e[0] == '0'); }();
  if (!use_graph || prof || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
    if (int rc = run_ops(m, dev_in, B, dev_workspace, dev_out_cls, dev_out_box, st, prof)) return rc;
    vbt::count_launches(m->kernels_last_run);
    return VBT_OK;
  }
  const vbt_model::GraphKey key{dev_in, dev_workspace, dev_out_cls, dev_out_box, B};
  auto it = m->graphs.find(key);
  if (it == m->graphs.end()) {
    auto seen = m->graph_seen.find(key);
    if (seen == m->graph_seen.end()) {
      // first call with these buffers: run directly (also performs every one-time
      // cudaFuncSetAttribute outside of a capture); capture on the next call
      m->graph_seen.insert(key);
      if (int rc = run_ops(m, dev_in, B, dev_workspace, dev_out_cls, dev_out_box, st, nullptr)) return rc;
      vbt::count_launches(m->kernels_last_run);
      return VBT_OK;
    }
    cudaGraph_t graph = nullptr;
    VBT_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = run_ops(m, dev_in, B, dev_workspace, dev_out_cls, dev_out_box, st, nullptr);
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != VBT_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    VBT_CHECK_CUDA(e);
    cudaGraphExec_t exec = nullptr;
    VBT_CHECK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    it = m->graphs.emplace(key, exec).first;
    m->graph_kernels[key] = m->kernels_last_run;
  }
  VBT_CHECK_CUDA(cudaGraphLaunch(it->second, st));
  vbt::count_launches(m->graph_kernels[key]);
  return VBT_OK;
}
