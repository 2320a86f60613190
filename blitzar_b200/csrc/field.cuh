// Prime-field arithmetic on 32-bit limbs for the four curves of the MSM hot path.
//
// Replaces (re-derived, not translated): sxt/field51/operation/{mul,sq,add,sub}.cc (radix-2^51
// curve25519 field), sxt/field12 (bls12-381), sxt/field25 (bn254), sxt/fieldgk (grumpkin) and
// sxt/base/field/arithmetic_utility.h:39-75 (mac/adc/sbb helpers).
//
// Representation on device:
//   * F25519: 8 x u32, plain residue, kept only loosely reduced (any value < 2^256 that is
//     congruent mod p = 2^255-19); products fold with 2^256 = 38 (mod p).
//   * Mont<P>: N x u32 Montgomery residues with R = 2^(32N). Because R equals the reference's
//     R (2^256 resp. 2^384, 64-bit limbs), the reference's in-memory Montgomery limbs are
//     bit-identical to ours: u64[N/2] little-endian == u32[N].
//
// All functions are __host__ __device__ so the same code is exercised by the CPU-side emulation
// tests (tests/emul); device builds replace the inner loops by carry-chain PTX.
#pragma once
#include "constants.cuh"

#define B200_HD __host__ __device__ __forceinline__

namespace b200 {

template <int N> struct alignas(16) Fe {
  u32 l[N];
};

// ------------------------------------------------------------------------------------------------
// multi-limb add / sub
// ------------------------------------------------------------------------------------------------
template <int N> B200_HD u32 limbs_add(u32* r, const u32* a, const u32* b) {
#ifdef __CUDA_ARCH__
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int i = 1; i < N; ++i)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
  u32 c;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
#else
  u64 c = 0;
  for (int i = 0; i < N; ++i) {
    c += (u64)a[i] + b[i];
    r[i] = (u32)c;
    c >>= 32;
  }
  return (u32)c;
#endif
}

// returns borrow (0 or 1)
template <int N> B200_HD u32 limbs_sub(u32* r, const u32* a, const u32* b) {
#ifdef __CUDA_ARCH__
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int i = 1; i < N; ++i)
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(a[i]), "r"(b[i]));
  u32 bw;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(bw));
  return bw & 1u;
#else
  u64 bw = 0;
  for (int i = 0; i < N; ++i) {
    u64 t = (u64)a[i] - b[i] - bw;
    r[i] = (u32)t;
    bw = t >> 63;
  }
  return (u32)bw;
#endif
}

// r = a + k (small), returns carry
template <int N> B200_HD u32 limbs_add_small(u32* r, const u32* a, u32 k) {
  u64 c = k;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    c += a[i];
    r[i] = (u32)c;
    c >>= 32;
  }
  return (u32)c;
}

template <int N> B200_HD u32 limbs_sub_small(u32* r, const u32* a, u32 k) {
  u64 bw = k;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    u64 t = (u64)a[i] - bw;
    r[i] = (u32)t;
    bw = t >> 63;
  }
  return (u32)bw;
}

template <int N> B200_HD bool limbs_is_zero(const u32* a) {
  u32 x = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
    x |= a[i];
  return x == 0;
}

// schoolbook product t[2N] = a * b
template <int N> B200_HD void limbs_mul_wide(u32* t, const u32* a, const u32* b) {
#pragma unroll
  for (int i = 0; i < 2 * N; ++i)
    t[i] = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      c += (u64)a[j] * b[i] + t[i + j];
      t[i + j] = (u32)c;
      c >>= 32;
    }
    t[i + N] = (u32)c;
  }
}

// ------------------------------------------------------------------------------------------------
// F25519: GF(2^255-19), loosely reduced 8-limb residues
// ------------------------------------------------------------------------------------------------
struct F25519 {
  static constexpr int N = 8;
  typedef Fe<8> E;

  static B200_HD E zero() {
    E r;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.l[i] = 0;
    return r;
  }
  static B200_HD E one() {
    E r = zero();
    r.l[0] = 1;
    return r;
  }
  template <class C> static B200_HD E constant(C c) {
    E r;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.l[i] = c(i);
    return r;
  }

  static B200_HD void add(E& r, const E& a, const E& b) {
    u32 c = limbs_add<8>(r.l, a.l, b.l);
    c = limbs_add_small<8>(r.l, r.l, 38u * c);
    r.l[0] += 38u * c;  // after a second wrap the value is < 38, so this cannot carry
  }
  static B200_HD void sub(E& r, const E& a, const E& b) {
    u32 bw = limbs_sub<8>(r.l, a.l, b.l);
    bw = limbs_sub_small<8>(r.l, r.l, 38u * bw);
    r.l[0] -= 38u * bw;  // after a second wrap the value is >= 2^256-38, so this cannot borrow
  }
  static B200_HD void neg(E& r, const E& a) {
    E z = zero();
    sub(r, z, a);
  }
  static B200_HD void dbl(E& r, const E& a) { add(r, a, a); }

  // r = a*b. 16-limb product folded twice with 2^256 = 38.
  static B200_HD void mul(E& r, const E& a, const E& b) {
    u32 t[16];
    limbs_mul_wide<8>(t, a.l, b.l);
    fold(r, t);
  }
  static B200_HD void sqr(E& r, const E& a) { mul(r, a, a); }

  static B200_HD void fold(E& r, const u32* t) {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c += (u64)t[i + 8] * 38u + t[i];
      r.l[i] = (u32)c;
      c >>= 32;
    }
    // c < 39; fold the ninth limb
    u32 c2 = limbs_add_small<8>(r.l, r.l, (u32)c * 38u);
    r.l[0] += 38u * c2;
  }

  // canonical representative in [0, p)
  static B200_HD void canonical(E& r, const E& a) {
    E v = a;
    u32 top = v.l[7] >> 31;
    v.l[7] &= 0x7fffffffu;
    limbs_add_small<8>(v.l, v.l, 19u * top);  // < 2^255 + 19
    E t;
    limbs_add_small<8>(t.l, v.l, 19u);  // v + 19 >= 2^255  <=>  v >= p
    u32 ge = t.l[7] >> 31;
    t.l[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.l[i] = ge ? t.l[i] : v.l[i];
  }
  static B200_HD bool is_zero(const E& a) {
    E c;
    canonical(c, a);
    return limbs_is_zero<8>(c.l);
  }
  static B200_HD bool is_negative(const E& a) {  // f51p::is_negative: lsb of canonical form
    E c;
    canonical(c, a);
    return c.l[0] & 1u;
  }
  static B200_HD bool equal(const E& a, const E& b) {
    E d;
    sub(d, a, b);
    return is_zero(d);
  }
  static B200_HD void select(E& r, const E& a, const E& b, bool pick_b) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.l[i] = pick_b ? b.l[i] : a.l[i];
  }
  static B200_HD void abs(E& r, const E& a) {
    E n;
    neg(n, a);
    select(r, a, n, is_negative(a));
  }

  // a^e for a public exponent given as 8 limbs
  template <class C> static B200_HD void pow(E& r, const E& a, C expo) {
    E acc = one();
    for (int i = 255; i >= 0; --i) {
      sqr(acc, acc);
      if ((expo(i >> 5) >> (i & 31)) & 1u)
        mul(acc, acc, a);
    }
    r = acc;
  }
  struct ExpPm2 {
    B200_HD u32 operator()(int i) const { return i == 0 ? 0xffffffebu : (i == 7 ? 0x7fffffffu : 0xffffffffu); }
  };
  struct ExpP58 {  // (p-5)/8 = 2^252 - 3
    B200_HD u32 operator()(int i) const { return i == 0 ? 0xfffffffdu : (i == 7 ? 0x0fffffffu : 0xffffffffu); }
  };
  static B200_HD void invert(E& r, const E& a) { pow(r, a, ExpPm2{}); }
  static B200_HD void pow22523(E& r, const E& a) { pow(r, a, ExpP58{}); }

  // radix-2^51 limbs (sxt_ristretto255 / c21t::element_p3 field layout; limbs may be unreduced)
  static B200_HD void from_radix51(E& r, const u64* h) {
    u32 acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i)
      acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int o = 51 * i, w = o >> 5, s = o & 31;
      u32 p0 = (u32)(h[i] << s);
      u32 p1 = (u32)((h[i] >> (32 - s)));
      u32 p2 = s ? (u32)(h[i] >> (64 - s)) : 0u;
      if (s == 0) {
        p0 = (u32)h[i];
        p1 = (u32)(h[i] >> 32);
      }
      u64 c = (u64)acc[w] + p0;
      acc[w] = (u32)c;
      c >>= 32;
      c += (u64)acc[w + 1] + p1;
      acc[w + 1] = (u32)c;
      c >>= 32;
      c += (u64)acc[w + 2] + p2;
      acc[w + 2] = (u32)c;
      c >>= 32;
#pragma unroll
      for (int k = w + 3; k < 10; ++k) {
        c += acc[k];
        acc[k] = (u32)c;
        c >>= 32;
      }
    }
    // value < 2^269: fold limbs 8,9 with 2^256 = 38
    u64 c = 0;
    u32 hi[8] = {acc[8], acc[9], 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c += (u64)hi[i] * 38u + acc[i];
      r.l[i] = (u32)c;
      c >>= 32;
    }
    u32 c2 = limbs_add_small<8>(r.l, r.l, (u32)c * 38u);
    r.l[0] += 38u * c2;
  }
  static B200_HD void to_radix51(u64* h, const E& a) {
    E c;
    canonical(c, a);
    u64 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = (u64)c.l[2 * i] | ((u64)c.l[2 * i + 1] << 32);
    const u64 m = 0x7ffffffffffffULL;
    h[0] = w[0] & m;
    h[1] = ((w[0] >> 51) | (w[1] << 13)) & m;
    h[2] = ((w[1] >> 38) | (w[2] << 26)) & m;
    h[3] = ((w[2] >> 25) | (w[3] << 39)) & m;
    h[4] = (w[3] >> 12) & m;
  }
  // 32 little-endian bytes, top bit ignored (f51b::from_bytes semantics)
  static B200_HD void from_bytes(E& r, const unsigned char* s) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r.l[i] = (u32)s[4 * i] | ((u32)s[4 * i + 1] << 8) | ((u32)s[4 * i + 2] << 16) |
               ((u32)s[4 * i + 3] << 24);
    r.l[7] &= 0x7fffffffu;
  }
  static B200_HD void to_bytes(unsigned char* s, const E& a) {
    E c;
    canonical(c, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[4 * i] = (unsigned char)c.l[i];
      s[4 * i + 1] = (unsigned char)(c.l[i] >> 8);
      s[4 * i + 2] = (unsigned char)(c.l[i] >> 16);
      s[4 * i + 3] = (unsigned char)(c.l[i] >> 24);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Mont<P>: Montgomery residues, fully reduced. P supplies N, p(i), one(i), r2(i), inv.
// ------------------------------------------------------------------------------------------------
template <class P> struct Mont {
  static constexpr int N = P::N;
  typedef Fe<N> E;

  static B200_HD E zero() {
    E r;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = 0;
    return r;
  }
  static B200_HD E one() {
    E r;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = P::one(i);
    return r;
  }
  static B200_HD E modulus() {
    E r;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = P::p(i);
    return r;
  }
  template <class C> static B200_HD E constant(C c) {
    E r;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = c(i);
    return r;
  }

  // all moduli leave >= 2 spare bits in N limbs, so a+b never carries out
  static B200_HD void add(E& r, const E& a, const E& b) {
    E s, d, p = modulus();
    limbs_add<N>(s.l, a.l, b.l);
    u32 bw = limbs_sub<N>(d.l, s.l, p.l);
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = bw ? s.l[i] : d.l[i];
  }
  static B200_HD void sub(E& r, const E& a, const E& b) {
    E d, e, p = modulus();
    u32 bw = limbs_sub<N>(d.l, a.l, b.l);
    limbs_add<N>(e.l, d.l, p.l);
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = bw ? e.l[i] : d.l[i];
  }
  static B200_HD void neg(E& r, const E& a) {
    E z = zero();
    sub(r, z, a);
  }
  static B200_HD void dbl(E& r, const E& a) { add(r, a, a); }
  static B200_HD bool is_zero(const E& a) { return limbs_is_zero<N>(a.l); }
  static B200_HD bool equal(const E& a, const E& b) {
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
      x |= a.l[i] ^ b.l[i];
    return x == 0;
  }
  static B200_HD void select(E& r, const E& a, const E& b, bool pick_b) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = pick_b ? b.l[i] : a.l[i];
  }

  // CIOS Montgomery product: r = a*b/R mod p
  static B200_HD void mul(E& r, const E& a, const E& b) {
    u32 t[N + 2];
#pragma unroll
    for (int i = 0; i < N + 2; ++i)
      t[i] = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      u64 c = 0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        c += (u64)a.l[j] * b.l[i] + t[j];
        t[j] = (u32)c;
        c >>= 32;
      }
      c += t[N];
      t[N] = (u32)c;
      t[N + 1] = (u32)(c >> 32);
      u32 m = t[0] * P::inv;
      c = ((u64)m * P::p(0) + t[0]) >> 32;
#pragma unroll
      for (int j = 1; j < N; ++j) {
        c += (u64)m * P::p(j) + t[j];
        t[j - 1] = (u32)c;
        c >>= 32;
      }
      c += t[N];
      t[N - 1] = (u32)c;
      t[N] = t[N + 1] + (u32)(c >> 32);
    }
    E s, d, p = modulus();
#pragma unroll
    for (int i = 0; i < N; ++i)
      s.l[i] = t[i];
    u32 bw = limbs_sub<N>(d.l, s.l, p.l);
    bool keep = bw && (t[N] == 0);
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = keep ? s.l[i] : d.l[i];
  }
  static B200_HD void sqr(E& r, const E& a) { mul(r, a, a); }

  template <class C> static B200_HD void pow(E& r, const E& a, C expo) {
    E acc = one();
    for (int i = 32 * N - 1; i >= 0; --i) {
      sqr(acc, acc);
      if ((expo(i >> 5) >> (i & 31)) & 1u)
        mul(acc, acc, a);
    }
    r = acc;
  }
  struct ExpPm2 {
    B200_HD u32 operator()(int i) const { return P::pm2(i); }
  };
  // r = 1/a (0 -> 0), Montgomery domain preserved
  static B200_HD void invert(E& r, const E& a) { pow(r, a, ExpPm2{}); }

  static B200_HD void from_mont(E& r, const E& a) {
    E o = zero();
    o.l[0] = 1;
    mul(r, a, o);
  }
  static B200_HD void to_mont(E& r, const E& a) {
    E r2;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r2.l[i] = P::r2(i);
    mul(r, a, r2);
  }
  // f12p::lexicographically_largest: plain value > (p-1)/2
  static B200_HD bool lexicographically_largest(const E& a) {
    E v, h, d;
    from_mont(v, a);
#pragma unroll
    for (int i = 0; i < N; ++i)
      h.l[i] = P::half(i);
    u32 bw = limbs_sub<N>(d.l, h.l, v.l);  // borrow <=> v > half
    return bw != 0;
  }
  // reference memory layout: u64 limbs little-endian == u32 limbs
  static B200_HD void load(E& r, const void* src) {
    const u32* s = (const u32*)src;
#pragma unroll
    for (int i = 0; i < N; ++i)
      r.l[i] = s[i];
  }
  static B200_HD void store(void* dst, const E& a) {
    u32* d = (u32*)dst;
#pragma unroll
    for (int i = 0; i < N; ++i)
      d[i] = a.l[i];
  }
};

struct BnParams {
  static constexpr int N = 8;
  static constexpr u32 inv = BN_INV;
  static B200_HD u32 p(int i) { return BN_P(i); }
  static B200_HD u32 one(int i) { return BN_ONE(i); }
  static B200_HD u32 r2(int i) { return BN_R2(i); }
  static B200_HD u32 pm2(int i) { return BN_PM2(i); }
  static B200_HD u32 half(int i) { return BN_HALF(i); }
};
struct GkParams {
  static constexpr int N = 8;
  static constexpr u32 inv = GK_INV;
  static B200_HD u32 p(int i) { return GK_P(i); }
  static B200_HD u32 one(int i) { return GK_ONE(i); }
  static B200_HD u32 r2(int i) { return GK_R2(i); }
  static B200_HD u32 pm2(int i) { return GK_PM2(i); }
  static B200_HD u32 half(int i) { return GK_HALF(i); }
};
struct BlsParams {
  static constexpr int N = 12;
  static constexpr u32 inv = BLS_INV;
  static B200_HD u32 p(int i) { return BLS_P(i); }
  static B200_HD u32 one(int i) { return BLS_ONE(i); }
  static B200_HD u32 r2(int i) { return BLS_R2(i); }
  static B200_HD u32 pm2(int i) { return BLS_PM2(i); }
  static B200_HD u32 half(int i) { return BLS_HALF(i); }
};
typedef Mont<BnParams> FBn;
typedef Mont<GkParams> FGk;
typedef Mont<BlsParams> FBls;

}  // namespace b200
