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
  __host__ __device__ Requant() {}
  __host__ __device__ Requant(int zp, int lo, int hi)
      : flo((float)(lo - zp)), fhi((float)(hi - zp)), magic_zp(kRoundMagicBits - zp) {}
  __device__ __forceinline__ int operator()(int acc, float mult) const {
    float y = __fmul_rn(__int2float_rn(acc), mult);
    y = fminf(fmaxf(y, flo), fhi);
    return __float_as_int(__fadd_rn(y, kRoundMagic)) - magic_zp;
  }
};

// four values already inside [-128, 127] -> one little-endian word of int8
__device__ __forceinline__ uint32_t pack4_s8(int y0, int y1, int y2, int y3) {
  const uint32_t lo = __byte_perm((uint32_t)y0, (uint32_t)y1, 0x0040);
  const uint32_t hi = __byte_perm((uint32_t)y2, (uint32_t)y3, 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

}  // namespace vbt
