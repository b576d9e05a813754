// Requantisation shared by the int8 kernels:  y = clamp(rne(float32(acc) * M) + zp, lo, hi).
//
// Same value as the straightforward int -> float -> multiply -> round -> add -> clamp
// sequence (oracle/effdet.py:_requant), reordered so that it costs one conversion
// instead of two and no integer min/max:
//   * the clamp is applied to the float product against (lo - zp, hi - zp): both bounds are
//     integers and rounding is monotonic, so clamp-then-round == round-then-clamp;
//   * the clamped product lies in [-256, 255], where adding 1.5 * 2^23 rounds it to the
//     nearest integer, ties to even -- exactly rint() -- and leaves that integer in the low
//     mantissa bits; subtracting (bits(1.5 * 2^23) - zp) yields rint(product) + zp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vbt {

constexpr float kRoundMagic = 12582912.0f;        // 1.5 * 2^23
constexpr int kRoundMagicBits = 0x4B400000;

struct Requant {
  float flo, fhi;     // float(lo - zp), float(hi - zp)
  int magic_zp;       // kRoundMagicBits - zp
  // packed path (pack4): the rounded products of two channels ride in one register as signed
  // 16-bit halves, so the clamp is one VIMNMX.S16x2 pair per two channels and the zero point
  // one 32-bit add per two channels.  Needs |acc * mult| < 2^15 - 256 for every input the op can
  // see; the model builder proves that per op (effdet.pack_blob -> OpRecord.requant_fast).
  uint32_t lo2, hi2;  // clamp bounds of rint(p) + 2^15 in both 16-bit halves (unsigned)
  uint32_t zp2;       // zp * 0x10001: one 32-bit add puts zp on both halves (biased halves stay
                      // inside (0, 2^16), so nothing carries from the low half into the high one)
  int fast;
  __host__ __device__ Requant() {}
  __host__ __device__ Requant(int zp, int lo, int hi, int fast_ok = 0)
      : flo((float)(lo - zp)), fhi((float)(hi - zp)), magic_zp(kRoundMagicBits - zp),
        lo2((uint32_t)(lo - zp + 32768) * 0x10001u), hi2((uint32_t)(hi - zp + 32768) * 0x10001u),
        zp2((uint32_t)zp * 0x10001u), fast(fast_ok) {}
  __device__ __forceinline__ int operator()(int acc, float mult) const {
    float y = __fmul_rn(__int2float_rn(acc), mult);
    y = fminf(fmaxf(y, flo), fhi);
    return __float_as_int(__fadd_rn(y, kRoundMagic)) - magic_zp;
  }
  // four accumulators (bias included) of consecutive channels -> one word of four int8.
  // Same value as four operator() calls: rounding is monotonic and the bounds are integers, so
  // clamp(rint(p)) == rint(clamp(p)); the magic add rounds p exactly like rint() (the zero point
  // is added afterwards, in integers, so ties still go to the even multiple of the output step).
  __device__ __forceinline__ uint32_t pack4(int a0, int a1, int a2, int a3, float m0, float m1, float m2,
                                            float m3) const;
  // the same with the choice of path made at compile time (kernels templated on it: no branch,
  // no divergence bookkeeping around every group of four channels)
  template <bool FAST>
  __device__ __forceinline__ uint32_t pack4t(int a0, int a1, int a2, int a3, float m0, float m1, float m2,
                                             float m3) const;
};

// four values already inside [-128, 127] -> one little-endian word of int8
__device__ __forceinline__ uint32_t pack4_s8(int y0, int y1, int y2, int y3) {
  const uint32_t lo = __byte_perm((uint32_t)y0, (uint32_t)y1, 0x0040);
  const uint32_t hi = __byte_perm((uint32_t)y2, (uint32_t)y3, 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

}  // namespace vbt

namespace vbt {

__device__ __forceinline__ uint32_t Requant::pack4(int a0, int a1, int a2, int a3, float m0, float m1,
                                                   float m2, float m3) const {
  return fast ? pack4t<true>(a0, a1, a2, a3, m0, m1, m2, m3) : pack4t<false>(a0, a1, a2, a3, m0, m1, m2, m3);
}

template <bool FAST>
__device__ __forceinline__ uint32_t Requant::pack4t(int a0, int a1, int a2, int a3, float m0, float m1,
                                                    float m2, float m3) const {
  if (!FAST) return pack4_s8((*this)(a0, m0), (*this)(a1, m1), (*this)(a2, m2), (*this)(a3, m3));
  // scalar multiplies + packed adds: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even with
  // --fmad=false, which would round once instead of twice
  const float p0 = __fmul_rn(__int2float_rn(a0), m0), p1 = __fmul_rn(__int2float_rn(a1), m1);
  const float p2 = __fmul_rn(__int2float_rn(a2), m2), p3 = __fmul_rn(__int2float_rn(a3), m3);
  unsigned long long q01, q23, mg;
  asm("mov.b64 %0, {%1, %2};" : "=l"(q01) : "f"(p0), "f"(p1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(q23) : "f"(p2), "f"(p3));
  asm("mov.b64 %0, {%1, %1};" : "=l"(mg) : "f"(kRoundMagic + 32768.0f));   // even: ties as rint()
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q01) : "l"(mg));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q23) : "l"(mg));
  uint32_t b0, b1, b2, b3;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(b0), "=r"(b1) : "l"(q01));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(b2), "=r"(b3) : "l"(q23));
  // low 16 bits of 0x4B400000 + 2^15 + rint(p) = rint(p) + 2^15, an unsigned half
  uint32_t h01 = __byte_perm(b0, b1, 0x5410), h23 = __byte_perm(b2, b3, 0x5410);
  h01 = __vminu2(__vmaxu2(h01, lo2), hi2) + zp2;
  h23 = __vminu2(__vmaxu2(h23, lo2), hi2) + zp2;
  return __byte_perm(h01, h23, 0x6420);                  // 2^15 + value: the low byte is the int8
}

}  // namespace vbt
