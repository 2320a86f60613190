// Warp-cooperative ("lane-sliced") arithmetic in GF(2^255 - 19) for the latency-bound tail of an MSM.
//
// The tail kernels (240 Horner doublings over the window sums, the 250-squaring inverse square root
// of the ristretto encoding) are ONE dependent chain per output: a single warp cannot issue a
// 256-bit multiplication faster than its SM sub-partition's multiplier pipe allows (72 IMAD.WIDE at
// ~5.6 cycles each, tests/micro/latency.cu), however few lanes do useful work. Here one field
// element is spread over 8 lanes — lane 8g + m of the warp holds limb m of the element of group g —
// so a multiplication is 16 multiply-adds per lane (two columns of the schoolbook product, operands
// fetched with warp shuffles) followed by a carry propagation ACROSS lanes: the high words move one
// and two lanes up by shuffle (times 38 where they wrap past 2^256) and the remaining single-bit
// carries are resolved for all lanes at once from two ballots (generate / propagate masks, one
// 8-bit addition = a carry-lookahead adder). Four elements (groups) are processed per warp, which is
// exactly the four independent products of one stage of the extended-coordinates point doubling.
//
// Replaces, for these chains, the per-thread schedules of field.cuh (which replace
// sxt/field51/operation/{mul,sq}.cc). Device only: the CPU emulation harness keeps the scalar path.
#pragma once
#include "curve.cuh"

#if defined(__CUDACC__) && !defined(B200_EMULATE)
namespace b200 {
namespace lane8 {

#define B200_DEV __device__ __forceinline__
constexpr unsigned kFull = 0xffffffffu;

B200_DEV u32 limb_index() { return threadIdx.x & 7u; }
B200_DEV u32 group_index() { return (threadIdx.x >> 3) & 3u; }

// Single-bit carries (generate flag cy per lane, limbs s) resolved over the 8 lanes of every group:
// the carry into limb i is bit i of (A + B) ^ A ^ B with A = propagate | generate, B = generate.
B200_DEV u32 resolve_carry(u32 s, u32 cy, u32& carry_out) {
  const u32 sh = threadIdx.x & 24u, m = threadIdx.x & 7u;
  const u32 G = (__ballot_sync(kFull, cy != 0) >> sh) & 0xffu;
  const u32 P = (__ballot_sync(kFull, s == 0xffffffffu) >> sh) & 0xffu;
  const u32 A = P | G, S = A + G;
  carry_out = (S >> 8) & 1u;
  return s + (((S ^ A ^ G) >> m) & 1u);
}
// borrows: generate = this limb borrowed, propagate = limb is zero
B200_DEV u32 resolve_borrow(u32 d, u32 bw, u32& borrow_out) {
  const u32 sh = threadIdx.x & 24u, m = threadIdx.x & 7u;
  const u32 G = (__ballot_sync(kFull, bw != 0) >> sh) & 0xffu;
  const u32 P = (__ballot_sync(kFull, d == 0u) >> sh) & 0xffu;
  const u32 A = P | G, S = A + G;
  borrow_out = (S >> 8) & 1u;
  return d - (((S ^ A ^ G) >> m) & 1u);
}
// value + carry_out * 2^256 == value + 38 * carry_out (mod p); at most two wraps can happen
B200_DEV u32 fold_carry(u32 r, u32 carry_out) {
  const u32 m = threadIdx.x & 7u;
  const u32 s = r + (m == 0 ? 38u * carry_out : 0u);
  u32 c2;
  u32 t = resolve_carry(s, s < r ? 1u : 0u, c2);
  return t + (m == 0 ? 38u * c2 : 0u);  // after a second wrap the value is < 2^7: cannot carry
}
B200_DEV u32 fold_borrow(u32 r, u32 borrow_out) {
  const u32 m = threadIdx.x & 7u;
  const u32 sub = m == 0 ? 38u * borrow_out : 0u;
  const u32 d = r - sub;
  u32 b2;
  u32 t = resolve_borrow(d, r < sub ? 1u : 0u, b2);
  return t - (m == 0 ? 38u * b2 : 0u);  // a second wrap leaves a value >= 2^256 - 76: cannot borrow
}

B200_DEV u32 add(u32 a, u32 b) {
  const u32 s = a + b;
  u32 co;
  const u32 r = resolve_carry(s, s < a ? 1u : 0u, co);
  return fold_carry(r, co);
}
B200_DEV u32 sub(u32 a, u32 b) {
  const u32 d = a - b;
  u32 bo;
  const u32 r = resolve_borrow(d, a < b ? 1u : 0u, bo);
  return fold_borrow(r, bo);
}

// limb m of a * b mod p (loosely reduced: any representative below 2^256)
B200_DEV u32 mul(u32 a, u32 b) {
  const u32 m = threadIdx.x & 7u;
  u32 l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;  // columns m and m + 8 of the product
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const u32 ai = __shfl_sync(kFull, a, i, 8);
    const u32 bj = __shfl_sync(kFull, b, (m - (u32)i) & 7u, 8);
    const bool low = (u32)i <= m;  // a_i b_(m-i) belongs to column m, otherwise to column m + 8
    const u32 bl = low ? bj : 0u, bh = low ? 0u : bj;
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(l0), "+r"(l1), "+r"(l2)
        : "r"(ai), "r"(bl));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(h0), "+r"(h1), "+r"(h2)
        : "r"(ai), "r"(bh));
  }
  // column m + 8 folds onto column m with 2^256 = 38
  const u64 t0 = (u64)h0 * 38u + l0;
  const u64 t1 = (u64)h1 * 38u + l1 + (t0 >> 32);
  const u32 v0 = (u32)t0, v1 = (u32)t1, v2 = h2 * 38u + l2 + (u32)(t1 >> 32);
  // the words above bit 32 move one and two limbs up (times 38 where they wrap around)
  const u32 r1 = __shfl_sync(kFull, v1, (m - 1u) & 7u, 8);
  const u32 r2 = __shfl_sync(kFull, v2, (m - 2u) & 7u, 8);
  const u64 w = (u64)v0 + (u64)r1 * (m == 0 ? 38u : 1u) + (u64)r2 * (m < 2 ? 38u : 1u);
  const u32 lo = (u32)w;
  u32 hi = __shfl_sync(kFull, (u32)(w >> 32), (m - 1u) & 7u, 8);  // < 2^7
  hi = m == 0 ? hi * 38u : hi;
  const u32 s = lo + hi;
  u32 co;
  const u32 r = resolve_carry(s, s < lo ? 1u : 0u, co);
  return fold_carry(r, co);
}

// replicated element (all 8 limbs in every lane) <-> lane-sliced
B200_DEV u32 slice(const F25519::E& e) {
  const u32 m = threadIdx.x & 7u;
  u32 v = e.l[0];
#pragma unroll
  for (int k = 1; k < 8; ++k)
    v = m == (u32)k ? e.l[k] : v;
  return v;
}
B200_DEV void gather(F25519::E& e, u32 v) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    e.l[k] = __shfl_sync(kFull, v, k, 8);
}
// the element of group g, in every group
B200_DEV u32 from_group(u32 v, u32 g) { return __shfl_sync(kFull, v, 8u * g + (threadIdx.x & 7u)); }

B200_DEV u32 sqr_n(u32 a, int n) {
  for (int i = 0; i < n; ++i)
    a = mul(a, a);
  return a;
}
// a^((p-5)/8) = a^(2^252 - 3), the chain of F25519::pow22523
B200_DEV u32 pow22523(u32 a) {
  u32 t0 = mul(a, a);            // 2
  u32 t1 = sqr_n(t0, 2);         // 8
  t1 = mul(a, t1);               // 9
  t0 = mul(t0, t1);              // 11
  u32 t2 = mul(t0, t0);          // 22
  t1 = mul(t1, t2);              // 2^5 - 1
  t2 = sqr_n(t1, 5);
  t1 = mul(t2, t1);              // 2^10 - 1
  t2 = sqr_n(t1, 10);
  t2 = mul(t2, t1);              // 2^20 - 1
  u32 t3 = sqr_n(t2, 20);
  t2 = mul(t3, t2);              // 2^40 - 1
  t2 = sqr_n(t2, 10);
  t1 = mul(t2, t1);              // 2^50 - 1
  t2 = sqr_n(t1, 50);
  t2 = mul(t2, t1);              // 2^100 - 1
  t3 = sqr_n(t2, 100);
  t2 = mul(t3, t2);              // 2^200 - 1
  t2 = sqr_n(t2, 50);
  t1 = mul(t2, t1);              // 2^250 - 1
  t1 = sqr_n(t1, 2);
  return mul(t1, a);
}

// Extended-coordinates point, one coordinate per group: X, Y, Z, T in groups 0..3.
// n doublings (the dbl-2008-hwcd schedule of Ed25519::dbl): stage 1 squares X, Y, Z, X + Y in the
// four groups at once, stage 2 multiplies E F, G H, F G, E H.
B200_DEV u32 dbl_n(u32 v, int n) {
  const u32 g = (threadIdx.x >> 3) & 3u;
  for (int it = 0; it < n; ++it) {
    const u32 X = from_group(v, 0), Y = from_group(v, 1);
    const u32 xy = add(X, Y);
    const u32 opnd = g == 3 ? xy : v;
    const u32 sq = mul(opnd, opnd);
    const u32 A = from_group(sq, 0), B = from_group(sq, 1), Cz = from_group(sq, 2),
              t1 = from_group(sq, 3);
    const u32 H = add(A, B), E = sub(H, t1), G = sub(A, B), C2 = add(Cz, Cz), Fv = add(C2, G);
    const u32 lhs = (g == 1) ? G : (g == 2 ? Fv : E);   // X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
    const u32 rhs = (g == 0) ? Fv : (g == 2 ? G : H);
    v = mul(lhs, rhs);
  }
  return v;
}
B200_DEV u32 slice_point(const Ed25519::Point& p) {
  const u32 g = (threadIdx.x >> 3) & 3u;
  const u32 x = slice(p.X), y = slice(p.Y), z = slice(p.Z), t = slice(p.T);
  return g == 0 ? x : (g == 1 ? y : (g == 2 ? z : t));
}
B200_DEV void gather_point(Ed25519::Point& p, u32 v) {
  const u32 m = threadIdx.x & 7u;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    p.X.l[k] = __shfl_sync(kFull, v, k);
    p.Y.l[k] = __shfl_sync(kFull, v, 8 + k);
    p.Z.l[k] = __shfl_sync(kFull, v, 16 + k);
    p.T.l[k] = __shfl_sync(kFull, v, 24 + k);
  }
  (void)m;
}
#undef B200_DEV

// Self-test body (b200_selftest_lane_arithmetic): every group of 8 lanes checks the lane-sliced
// operations against the per-thread schedules of field.cuh / curve.cuh on pseudo-random and edge-case
// operands; mismatches are counted.
struct SelfTestBody {
  static constexpr int kBlock = 32;
  u32 seed;
  u32* mismatches;
  static __device__ u32 rnd(u64& st) {
    st ^= st << 13;
    st ^= st >> 7;
    st ^= st << 17;
    return (u32)(st >> 16);
  }
  static __device__ bool same(const F25519::E& x, const F25519::E& y) {
    F25519::E a, b;
    F25519::canonical(a, x);
    F25519::canonical(b, y);
    bool ok = true;
    for (int i = 0; i < 8; ++i)
      ok = ok && a.l[i] == b.l[i];
    return ok;
  }
  __device__ void operator()(u64 tid) const {
    typedef F25519 F;
    const u32 group = (u32)(tid >> 3);
    u64 st = ((u64)seed << 32) ^ (0x9E3779B97F4A7C15ull * (group + 1));
    F::E a, b, want, got;
    for (int i = 0; i < 8; ++i) {
      a.l[i] = rnd(st);
      b.l[i] = rnd(st);
    }
    switch (group % 9) {  // edge cases: all-ones limbs, zero, p - 1, tiny, top bits
    case 1: for (int i = 0; i < 8; ++i) a.l[i] = 0xffffffffu; break;
    case 2: for (int i = 0; i < 8; ++i) a.l[i] = b.l[i] = 0xffffffffu; break;
    case 3: for (int i = 0; i < 8; ++i) b.l[i] = 0; break;
    case 4: for (int i = 0; i < 8; ++i) a.l[i] = 0xffffffffu; a.l[0] = 0xffffffecu; a.l[7] = 0x7fffffffu; break;
    case 5: for (int i = 1; i < 8; ++i) a.l[i] = 0; a.l[0] = 37; break;
    case 6: for (int i = 0; i < 7; ++i) b.l[i] = 0xffffffffu; break;
    case 7: for (int i = 0; i < 8; ++i) a.l[i] = 0; for (int i = 0; i < 8; ++i) b.l[i] = 0xffffffffu; break;
    default: break;
    }
    u32 bad = 0;
    const u32 la = lane8::slice(a), lb = lane8::slice(b);
    lane8::gather(got, lane8::mul(la, lb));
    F::mul(want, a, b);
    bad += same(got, want) ? 0 : 1;
    lane8::gather(got, lane8::add(la, lb));
    F::add(want, a, b);
    bad += same(got, want) ? 0 : 1;
    lane8::gather(got, lane8::sub(la, lb));
    F::sub(want, a, b);
    bad += same(got, want) ? 0 : 1;
    lane8::gather(got, lane8::sub(lb, la));
    F::sub(want, b, a);
    bad += same(got, want) ? 0 : 1;
    lane8::gather(got, lane8::mul(lane8::sub(la, lb), lane8::add(la, la)));
    {
      F::E d, s2;
      F::sub(d, a, b);
      F::add(s2, a, a);
      F::mul(want, d, s2);
    }
    bad += same(got, want) ? 0 : 1;
    if (group % 4 == 0) {  // the long chain, on a quarter of the groups
      lane8::gather(got, lane8::pow22523(la));
      F::pow22523(want, a);
      bad += same(got, want) ? 0 : 1;
    }
    // point doubling: the whole warp holds ONE point (the built-in generator of this warp)
    Ed25519::Point p, q;
    Ed25519::builtin_generator(p, (u64)seed * 1000u + (tid >> 5));
    q = p;
    for (int i = 0; i < 5; ++i)
      Ed25519::dbl(q, q);
    Ed25519::Point r;
    lane8::gather_point(r, lane8::dbl_n(lane8::slice_point(p), 5));
    bad += (same(r.X, q.X) && same(r.Y, q.Y) && same(r.Z, q.Z) && same(r.T, q.T)) ? 0 : 1;
    if (bad && (tid & 7u) == 0)
      atomicAdd(mismatches, bad);
  }
};

}  // namespace lane8
}  // namespace b200
#endif
