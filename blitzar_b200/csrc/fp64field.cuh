// GF(2^255-19) multiplication on the FP64 pipe, for the bucket-accumulation kernel only.
//
// Measured on B200 (tests/micro/pipes.cu): IMAD.WIDE.U32 issues at 23 lane-ops/clk/SM, DFMA at 61.
// A 52x52-bit product costs three FP64 instructions here (two DFMA.RZ + one DADD) against about
// six-and-a-half IMAD.WIDE for the same bits, so the field multiplication of the hot loop moves to
// double precision with EXACT integer semantics:
//
//   * a field element is five signed limbs, radix 2^51 ("balanced": |limb| <= 2^50 + 2^14), held as
//     int64; lazy add / sub are limb-wise on exact doubles (|operand limb| < 2^51.2);
//   * limb products p = a_i * b_j (|p| < 2^103) are split exactly with round-toward-zero FMAs:
//       h = fma_rz(a, b, 1.5*2^104)          = 1.5*2^104 + 2^52 * floor(p / 2^52)
//       l = fma_rz(a, b, (1.5*2^104 + 2^52) - h) = 2^52 + (p mod 2^52)
//     whose mantissas ARE the integers floor(p/2^52) and p mod 2^52, so the column sums are plain
//     int64 additions of the raw bit patterns (the exponent constants are subtracted once per column);
//   * columns 5..9 fold with 2^255 = 19, then one balanced carry chain renormalises.
//
// Host (emulation) builds evaluate fma_rz with exact 128-bit integer arithmetic, so the same
// schedule is verified bit-for-bit on a CPU (tests/emul: emul_check_fp64).
#pragma once
#include "field.cuh"

namespace b200 {

typedef long long i64;

struct FeD {
  i64 l[5];  // value = sum l[i] * 2^(51 i), balanced limbs
};

B200_HD double fma_rz_exact(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rz(a, b, c);
#else
  // every use in this file has integer-valued operands below 2^106: evaluate exactly, then
  // truncate toward zero to 53 significant bits
  __int128 s = (__int128)(i64)a * (__int128)(i64)b;
  // c may exceed 2^63: decompose c = ch * 2^52 + cl with integer parts
  double ch = __builtin_floor(c / 4503599627370496.0);
  double cl = c - ch * 4503599627370496.0;
  s += ((__int128)(i64)ch << 52) + (__int128)(i64)cl;
  bool neg = s < 0;
  unsigned __int128 m = neg ? (unsigned __int128)(-s) : (unsigned __int128)s;
  int bits = 0;
  for (unsigned __int128 t = m; t; t >>= 1)
    ++bits;
  if (bits > 53)
    m = (m >> (bits - 53)) << (bits - 53);  // truncate toward zero
  double r = 0;
  for (int i = 0; i < 128; i += 32)
    r += __builtin_ldexp((double)(unsigned)(m >> i), i);
  return neg ? -r : r;
#endif
}
B200_HD i64 double_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  i64 r;
  __builtin_memcpy(&r, &d, 8);
  return r;
#endif
}
B200_HD double bits_double(i64 v) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(v);
#else
  double r;
  __builtin_memcpy(&r, &v, 8);
  return r;
#endif
}

struct F25519D {
  static constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
  static constexpr i64 kMagicBits = 0x4338000000000000LL;
  static constexpr double kC1 = 30423614405477505635920876929024.0;  // 1.5 * 2^104
  static constexpr i64 kC1Bits = 0x4678000000000000LL;
  static constexpr double kC2 = 4503599627370496.0;               // 2^52
  static constexpr i64 kC2Bits = 0x4330000000000000LL;
  static constexpr double kC12 = 30423614405477510139520504299520.0;  // 1.5 * 2^104 + 2^52
  static constexpr i64 kMask51 = (1LL << 51) - 1;

  // exact int -> double for |x| < 2^51
  static B200_HD double to_double(i64 x) { return bits_double(x + kMagicBits) - kMagic; }

  // 8 x u32 canonical value (< 2^255) -> five unsigned 51-bit limbs
  static B200_HD void from_fe(FeD& r, const Fe<8>& a) {
    u64 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = (u64)a.l[2 * i] | ((u64)a.l[2 * i + 1] << 32);
    r.l[0] = (i64)(w[0] & (u64)kMask51);
    r.l[1] = (i64)(((w[0] >> 51) | (w[1] << 13)) & (u64)kMask51);
    r.l[2] = (i64)(((w[1] >> 38) | (w[2] << 26)) & (u64)kMask51);
    r.l[3] = (i64)(((w[2] >> 25) | (w[3] << 39)) & (u64)kMask51);
    r.l[4] = (i64)(w[3] >> 12);  // < 2^51 for canonical input (< 2^52 otherwise)
  }
  // balanced limbs -> loosely reduced 8 x u32 (adds 4p so every limb is non-negative)
  static B200_HD void to_fe(Fe<8>& r, const FeD& a) {
    u64 h[5];
    h[0] = (u64)(a.l[0] + 4 * ((1LL << 51) - 19));
#pragma unroll
    for (int i = 1; i < 5; ++i)
      h[i] = (u64)(a.l[i] + 4 * ((1LL << 51) - 1));
    F25519::from_radix51(r, h);
  }

  // r = a * b for exact-double operands with |limb| < 2^51.5; result balanced
  static B200_HD void mul(FeD& r, const double* a, const double* b) {
    u64 lo[9], hi[9];  // raw bit patterns, summed modulo 2^64
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      lo[k] = 0;
      hi[k] = 0;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double h = fma_rz_exact(a[i], b[j], kC1);
        double l = fma_rz_exact(a[i], b[j], kC12 - h);
        hi[i + j] += (u64)double_bits(h);
        lo[i + j] += (u64)double_bits(l);
      }
    }
    // strip the exponent constants: column k holds n_k products
    i64 R[10], L[9], H[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const u64 n = k < 5 ? k + 1 : 9 - k;
      L[k] = (i64)(lo[k] - n * (u64)kC2Bits);
      H[k] = (i64)(hi[k] - n * (u64)kC1Bits);
    }
    // 2^52 = 2 * 2^51: the high parts land one column up, doubled
    R[0] = L[0];
#pragma unroll
    for (int k = 1; k < 9; ++k)
      R[k] = L[k] + 2 * H[k - 1];
    R[9] = 2 * H[8];
    // fold columns 5..9 with 2^255 = 19
    i64 t[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
      t[k] = R[k] + 19 * R[k + 5];
    carry(r, t);
  }

  // balanced carry chain: limbs 1..4 end in [-2^50, 2^50), limb 0 within 2^50 + 2^14
  static B200_HD void carry(FeD& r, i64* t) {
    i64 c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c = (t[k] + (1LL << 50)) >> 51;
      t[k] -= c << 51;
      t[k + 1] += c;
    }
    c = (t[4] + (1LL << 50)) >> 51;
    t[4] -= c << 51;
    t[0] += 19 * c;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      r.l[k] = t[k];
  }
};

}  // namespace b200
