// Synthetic input generation on the device: the generators the reference's own benchmarks and tests
// use, produced at BASELINE sizes (2^22 bls12-381, 2^24 bn254 points) in a fraction of a second.
//
//   ristretto255 : built-in generator g(first + i)  (sqcgn::compute_base_element,
//                  sxt/seqcommit/generator/base_element.cc:30-35)
//   Weierstrass  : cg1rn / cn1rn / cgkrn::generate_random_element with
//                  basn::fast_random_number_generator{i + 1, i + 2}
//                  (sxt/curve_g1/random/element_p2.h:38-50, cbindings/pedersen.t.cc:81-123,
//                  benchmark/multi_exp_pip/benchmark.m.cc:84-95): four xorshift128+ outputs form a
//                  little-endian 32-byte scalar k_i, the point is (k_i mod 2^255) * G.
// The reference walks a 255-step double-and-add per point; here a 64 x 15 table of d * 16^j * G is
// built once (SynthTableBody) and every point is 64 mixed additions of table entries. Points come out
// distinct, match the reference's at every index (tests/test_gpu_baseline_sizes.py) and have known
// discrete logarithms k_i, which gives an exact full-size check of an MSM: sum_i s_i G_i =
// (sum_i s_i k_i mod r) * G.
#pragma once
#include "curve.cuh"
#include "runtime.cuh"

namespace b200 {

// 32 bytes of basn::fast_random_number_generator{index + 1, index + 2}
// (sxt/base/num/fast_random_number_generator.h:27-50), as 8 little-endian u32 words
B200_HD void synth_scalar_words(u32 k[8], u64 index) {
  u64 sa = index + 1, sb = index + 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    u64 t = sa, s = sb;
    sa = s;
    t ^= t << 23;
    t ^= t >> 17;
    t ^= s ^ (s >> 26);
    sb = t;
    u64 w = t + s;
    k[2 * j] = (u32)w;
    k[2 * j + 1] = (u32)(w >> 32);
  }
}

constexpr int kSynthWindows = 64, kSynthDigits = 15;  // 4-bit comb over 256 bits

// table[j * 15 + d - 1] = d * 16^j * G (affine)
template <class C> struct SynthTableBody {
  static constexpr int kBlock = 64;
  typename C::Gen* table;
  B200_HD void operator()(u64 t) const {
    const u32 j = (u32)(t / kSynthDigits), d = (u32)(t % kSynthDigits) + 1;
    typename C::Point base, acc = C::identity();
    C::gen_to_point(base, C::subgroup_generator(), false);
    for (u32 i = 0; i < 4 * j; ++i)
      C::dbl(base, base);
    for (int b = 3; b >= 0; --b) {
      C::dbl(acc, acc);
      if ((d >> b) & 1u)
        C::add(acc, acc, base);
    }
    typename C::Gen g;
    C::to_affine(g.x, g.y, acc);  // never the identity: d * 16^j < group order
    table[t] = g;
  }
};

template <class C> B200_HD void synth_point(typename C::Point& acc, const typename C::Gen* table,
                                            u64 index) {
  u32 k[8];
  synth_scalar_words(k, index);
  k[7] &= 0x7fffffffu;  // scalar_multiply255 skips the top bit (curve_g1/operation/scalar_multiply.cc:33-36)
  acc = C::identity();
  for (int j = 0; j < kSynthWindows; ++j) {
    const u32 d = (k[j >> 3] >> (4 * (j & 7))) & 15u;
    if (d)
      C::add_gen(acc, acc, table[j * kSynthDigits + d - 1], false);
  }
}

// kProjective: {X, Y, Z} structs (handle input); else affine {X, Y, infinity} at the commit stride
template <class C, bool kProjective> struct SynthGeneratorBody {
  static constexpr int kBlock = 128;
  const typename C::Gen* table;
  unsigned char* out;
  u64 first;
  B200_HD void operator()(u64 i) const {
    typename C::Point p;
    synth_point<C>(p, table, first + i);
    if (kProjective) {
      C::store_proj_abi(out + i * C::kAbiProjBytes, p);
    } else {
      typename C::fe x, y;
      const bool inf = C::to_affine(x, y, p);
      unsigned char* d = out + i * C::kAbiGenBytes;
      C::F::store(d, x);
      C::F::store(d + 4 * C::N, y);
      u32* tail = (u32*)(d + 8 * C::N);
      tail[0] = inf ? 1u : 0u;
      tail[1] = 0u;
    }
  }
};

// ristretto: built-in generators straight into the projective ABI layout, the exact (X:Y:Z:T) the
// derivation produces (no round trip through the cached form)
struct BuiltinGeneratorAbiBody {
  static constexpr int kBlock = 64;
  unsigned char* out;
  u64 first;
  B200_HD void operator()(u64 i) const {
    Ed25519::Point g;
    Ed25519::builtin_generator(g, first + i);
    Ed25519::store_proj_abi(out + i * Ed25519::kAbiProjBytes, g);
  }
};

template <class C> struct Synth {
  static void generators(stream_t s, void* out_dev, u64 n, u64 first, bool projective) {
    if constexpr (C::kCurveId == kRistretto255) {
      launch(BuiltinGeneratorAbiBody{(unsigned char*)out_dev, first}, n, s);
    } else {
      typedef typename C::Gen Gen;
      const u64 entries = (u64)kSynthWindows * kSynthDigits;
      Gen* table = (Gen*)dev_alloc(entries * sizeof(Gen), s);
      launch(SynthTableBody<C>{table}, entries, s);
      if (projective)
        launch(SynthGeneratorBody<C, true>{table, (unsigned char*)out_dev, first}, n, s);
      else
        launch(SynthGeneratorBody<C, false>{table, (unsigned char*)out_dev, first}, n, s);
      dev_free(table, s);
    }
  }
};

}  // namespace b200
