// Host-side arithmetic modulo the ristretto255 group order
//   l = 2^252 + 27742317777372353535851937790883648493
// for the inner-product argument (scalar folds, challenges, verification exponents). Replaces the
// parts of sxt/scalar25 (operation/{mul,muladd,add,sub,neg,inv,reduce}.cc) that the proof uses.
// Values are canonical 32-byte little-endian integers < l at every interface; products go through
// a 4 x 64-bit Montgomery multiplication.
#pragma once
#include <cstdint>
#include <cstring>

namespace b200 {

struct Sc {
  uint64_t v[4];
};

namespace sc_detail {
typedef unsigned __int128 u128;
constexpr uint64_t kL[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL};
inline bool geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] > b[i])
      return true;
    if (a[i] < b[i])
      return false;
  }
  return true;
}
inline uint64_t addn(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
inline uint64_t subn(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t bw = 0;
  for (int i = 0; i < 4; ++i) {
    u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 127);
  }
  return bw;
}
struct Consts {
  uint64_t inv;    // -l^-1 mod 2^64
  uint64_t r2[4];  // 2^512 mod l
  Consts() {
    uint64_t x = 1;
    for (int i = 0; i < 6; ++i)
      x *= 2 - kL[0] * x;
    inv = 0 - x;
    uint64_t t[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; ++i) {
      uint64_t c = addn(t, t, t);
      if (c || geq(t, kL))
        subn(t, t, kL);
    }
    std::memcpy(r2, t, 32);
  }
};
inline const Consts& consts() {
  static const Consts c;
  return c;
}
// a * b / 2^256 mod l for a * b < 2^256 * l
inline void mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  const Consts& K = consts();
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * K.inv;
    c = ((u128)m * kL[0] + t[0]) >> 64;
    for (int j = 1; j < 4; ++j) {
      c += (u128)m * kL[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || geq(t, kL))
    subn(t, t, kL);
  std::memcpy(r, t, 32);
}
}  // namespace sc_detail

inline Sc sc_zero() { return Sc{{0, 0, 0, 0}}; }
inline Sc sc_one() { return Sc{{1, 0, 0, 0}}; }
inline Sc sc_load(const uint8_t* bytes32) {
  Sc r;
  std::memcpy(r.v, bytes32, 32);
  return r;
}
inline void sc_store(uint8_t* bytes32, const Sc& a) { std::memcpy(bytes32, a.v, 32); }
// any 256-bit value -> canonical residue (s25o::reduce32)
inline Sc sc_reduce(const Sc& a) {
  using namespace sc_detail;
  Sc t, r;
  mont_mul(t.v, a.v, consts().r2);  // a * R mod l
  const uint64_t one[4] = {1, 0, 0, 0};
  mont_mul(r.v, t.v, one);
  return r;
}
inline Sc sc_mul(const Sc& a, const Sc& b) {
  using namespace sc_detail;
  Sc t, r;
  mont_mul(t.v, a.v, b.v);          // a b / R
  mont_mul(r.v, t.v, consts().r2);  // a b
  return r;
}
// Montgomery form x*R of a canonical x; sc_mul_mont(xm, y) = x*y with ONE Montgomery product
inline Sc sc_to_mont(const Sc& a) {
  using namespace sc_detail;
  Sc r;
  mont_mul(r.v, a.v, consts().r2);
  return r;
}
inline Sc sc_mul_mont(const Sc& a_mont, const Sc& b) {
  using namespace sc_detail;
  Sc r;
  mont_mul(r.v, a_mont.v, b.v);
  return r;
}
inline Sc sc_add(const Sc& a, const Sc& b) {
  using namespace sc_detail;
  Sc r;
  uint64_t c = addn(r.v, a.v, b.v);
  if (c || geq(r.v, kL))
    subn(r.v, r.v, kL);
  return r;
}
inline Sc sc_sub(const Sc& a, const Sc& b) {
  using namespace sc_detail;
  Sc r;
  if (subn(r.v, a.v, b.v))
    addn(r.v, r.v, kL);
  return r;
}
inline Sc sc_neg(const Sc& a) { return sc_sub(sc_zero(), a); }
inline Sc sc_muladd(const Sc& a, const Sc& b, const Sc& c) { return sc_add(sc_mul(a, b), c); }
// a^(l-2)
inline Sc sc_inv(const Sc& a) {
  using namespace sc_detail;
  uint64_t e[4];
  const uint64_t two[4] = {2, 0, 0, 0};
  subn(e, kL, two);
  Sc acc = sc_one();
  for (int i = 255; i >= 0; --i) {
    acc = sc_mul(acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1)
      acc = sc_mul(acc, a);
  }
  return acc;
}
// <a, b> over the first n entries of 32-byte scalars
inline Sc sc_inner_product(const uint8_t* a, const uint8_t* b, size_t n) {
  using namespace sc_detail;
  Sc acc = sc_zero();  // sum of a_i b_i / R; the factor R is restored once at the end
  for (size_t i = 0; i < n; ++i) {
    Sc t;
    mont_mul(t.v, sc_load(a + 32 * i).v, sc_load(b + 32 * i).v);
    acc = sc_add(acc, t);
  }
  Sc r;
  mont_mul(r.v, acc.v, consts().r2);
  return r;
}

}  // namespace b200
