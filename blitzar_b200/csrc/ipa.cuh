// Inner-product argument (Bulletproofs-style) over ristretto255 on top of the MSM engine:
// sxt_curve25519_prove_inner_product / sxt_curve25519_verify_inner_product.
//
// Replaces sxt/proof/inner_product/{proof_computation,gpu_driver,fold,generator_fold,
// verification_computation}.cc (+ generator_fold_kernel / scalar_fold_kernel). Protocol restated
// from cbindings/blitzar_api.h:479-611 and proof_computation.cc:61-155:
//   round j: L = <a_lo, G_hi> + <a_lo, b_hi> Q,  R = <a_hi, G_lo> + <a_hi, b_lo> Q
//            x  = transcript("L", L; "R", R; challenge "x")
//            a' = x a_lo + x^-1 a_hi,  b' = x^-1 b_lo + x b_hi,  G' = x^-1 G_lo + x G_hi
// Generators stay resident in HBM across rounds (folded in place by one kernel); the two MSMs of a
// round run through the Pippenger engine; scalar folds and the transcript are host work.
#pragma once
#include <vector>

#include "engine.cuh"
#include "transcript.h"

namespace b200 {

// G'[i] = m_lo * G[i] + m_hi * G[mid + i] by a shared-doubling (Shamir) ladder; the two scalars are
// the same for every thread, so the table index is warp-uniform.
struct FoldGeneratorsBody {
  static constexpr int kBlock = 64;
  const Ed25519::Gen* g;
  u32 mid;
  u32 m_lo[8], m_hi[8];
  Ed25519::Gen* out;
  B200_HD void operator()(u64 i) const {
    typedef Ed25519 C;
    C::Gen table[3];
    table[0] = g[i];
    table[1] = g[(u64)mid + i];
    C::Point p0, p1, p2;
    C::gen_to_point(p0, table[0], false);
    C::gen_to_point(p1, table[1], false);
    C::add(p2, p0, p1);
    C::point_to_gen(table[2], p2);
    C::Point acc = C::identity();
    for (int bit = 252; bit >= 0; --bit) {
      C::dbl(acc, acc);
      u32 sel = ((m_lo[bit >> 5] >> (bit & 31)) & 1u) | (((m_hi[bit >> 5] >> (bit & 31)) & 1u) << 1);
      if (sel)
        C::add_gen(acc, acc, table[sel - 1], false);
    }
    C::Gen r;
    C::point_to_gen(r, acc);
    out[i] = r;
  }
};
// compressed ristretto points -> device generator layout (invalid encodings become the identity
// and raise *bad)
struct DecodeToGenBody {
  static constexpr int kBlock = 32;
  const unsigned char* bytes;
  Ed25519::Gen* out;
  u32* bad;
  B200_HD void operator()(u64 i) const {
    Ed25519::Point p;
    if (!Ed25519::decode(p, bytes + 32 * i)) {
      p = Ed25519::identity();
      B200_ATOMIC_ADD(bad, 1u);
    }
    Ed25519::point_to_gen(out[i], p);
  }
};

struct Ipa {
  typedef Ed25519 C;
  typedef CurveOps<C> Ops;

  static unsigned ceil_log2(uint64_t n) {
    unsigned k = 0;
    while ((1ull << k) < n)
      ++k;
    return k;
  }
  static void scalar_words(u32 w[8], const Sc& s) {
    for (int i = 0; i < 4; ++i) {
      w[2 * i] = (u32)s.v[i];
      w[2 * i + 1] = (u32)(s.v[i] >> 32);
    }
  }
  // device generators g(offset .. offset+count): the precomputed table when it covers the range
  static const C::Gen* generators(const EngineCtx& ctx, DevBuf<C::Gen>& storage, uint64_t offset,
                                  uint64_t count) {
    if (offset + count <= ctx.num_builtin)
      return (const C::Gen*)ctx.builtin + offset;
    launch(BuiltinGeneratorBody{storage.p, offset}, count, ctx.s);
    return storage.p;
  }
  // sum_i s_i * gens[i] (+ extra_scalar * extra_gen) -> 32-byte ristretto encoding on the host
  static void msm_compressed(const EngineCtx& ctx, uint8_t* out32, const C::Gen* gens, uint64_t n,
                             const uint8_t* scalars, const C::Gen* extra_gen,
                             const Sc* extra_scalar) {
    stream_t s = ctx.s;
    const uint64_t total = n + (extra_gen ? 1 : 0);
    DevBuf<C::Gen> tg(total ? total : 1, s);
    DevBuf<unsigned char> ts(32 * total + 32, s);
    copy_d2d(tg.p, gens, n * sizeof(C::Gen), s);
    copy_h2d(ts.p, scalars, 32 * n, s);
    uint8_t extra[32];
    if (extra_gen) {
      copy_d2d(tg.p + n, extra_gen, sizeof(C::Gen), s);
      sc_store(extra, *extra_scalar);
      copy_h2d(ts.p + 32 * n, extra, 32, s);
    }
    std::vector<ColumnDesc> cols(1);
    cols[0].base = ts.p;
    cols[0].row_stride = 32;
    cols[0].bit_offset = 0;
    cols[0].bit_width = 256;
    cols[0].n = (u32)total;
    cols[0].is_signed = 0;
    cols[0].first_window = cols[0].num_windows = 0;
    DevBuf<C::Point> pt(1, s);
    DevBuf<unsigned char> enc(32, s);
    Ops::run_columns(ctx, tg.p, cols, pt.p);
    launch(StoreBody<C, true>{pt.p, enc.p}, 1, s);
    copy_d2h(out32, enc.p, 32, s);
    stream_sync(s);
  }
  // m_lo * v[i] + m_hi * v[mid + i], with v zero-padded to 2 * mid (prfip::fold_scalars)
  static std::vector<uint8_t> fold_scalars(const std::vector<uint8_t>& v, const Sc& m_lo,
                                           const Sc& m_hi, uint64_t mid) {
    const uint64_t len = v.size() / 32, p = len - mid;
    std::vector<uint8_t> r(32 * mid);
    const Sc lo_m = sc_to_mont(m_lo), hi_m = sc_to_mont(m_hi);  // one Montgomery product per term
    for (uint64_t i = 0; i < mid; ++i) {
      Sc t = sc_mul_mont(lo_m, sc_load(&v[32 * i]));
      if (i < p)
        t = sc_add(t, sc_mul_mont(hi_m, sc_load(&v[32 * (mid + i)])));
      sc_store(&r[32 * i], t);
    }
    return r;
  }
  static void init_transcript(Transcript& t, uint64_t n) {
    const char* domain = "inner product proof v1";
    t.append_message("domain-sep", (const uint8_t*)domain, std::strlen(domain));
    t.append_message("n", (const uint8_t*)&n, 8);
  }
  static Sc round_challenge(Transcript& t, const uint8_t* l32, const uint8_t* r32) {
    t.append_message("L", l32, 32);
    t.append_message("R", r32, 32);
    return t.challenge_scalar("x");
  }

  static void prove(const EngineCtx& ctx, uint8_t* l_vector, uint8_t* r_vector, uint8_t* ap_value,
                    uint8_t* transcript203, uint64_t n, uint64_t generators_offset,
                    const uint8_t* a_vector, const uint8_t* b_vector) {
    stream_t s = ctx.s;
    const unsigned k = ceil_log2(n);
    const uint64_t np = 1ull << k;
    Transcript tr(transcript203);
    init_transcript(tr, n);
    if (n == 1) {
      std::memcpy(ap_value, a_vector, 32);
      return;
    }
    DevBuf<C::Gen> gstore(np + 1, s);
    const C::Gen* G0 = generators(ctx, gstore, generators_offset, np + 1);
    const C::Gen* Q = G0 + np;
    DevBuf<C::Gen> gwork(np / 2, s);
    const C::Gen* G = G0;
    std::vector<uint8_t> a(a_vector, a_vector + 32 * n), b(b_vector, b_vector + 32 * n);
    uint64_t len = np;
    for (unsigned round = 0; round < k; ++round) {
      const uint64_t mid = len / 2;
      const uint64_t na = a.size() / 32, nb = b.size() / 32;
      const uint64_t a_hi = na - mid, b_hi = nb - mid;
      Sc c_l = sc_inner_product(&a[0], &b[32 * mid], std::min<uint64_t>(mid, b_hi));
      Sc c_r = sc_inner_product(&a[32 * mid], &b[0], std::min<uint64_t>(a_hi, mid));
      uint8_t* l_out = l_vector + 32 * round;
      uint8_t* r_out = r_vector + 32 * round;
      msm_compressed(ctx, l_out, G + mid, mid, &a[0], Q, &c_l);
      msm_compressed(ctx, r_out, G, a_hi, &a[32 * mid], Q, &c_r);
      Sc x = round_challenge(tr, l_out, r_out);
      Sc x_inv = sc_inv(x);
      a = fold_scalars(a, x, x_inv, mid);
      if (mid == 1)
        break;
      b = fold_scalars(b, x_inv, x, mid);
      FoldGeneratorsBody body;
      body.g = G;
      body.mid = (u32)mid;
      scalar_words(body.m_lo, x_inv);
      scalar_words(body.m_hi, x);
      body.out = gwork.p;
      launch(body, mid, s);
      G = gwork.p;
      len = mid;
    }
    stream_sync(s);
    std::memcpy(ap_value, &a[0], 32);
  }

  static int verify(const EngineCtx& ctx, uint8_t* transcript203, uint64_t n,
                    uint64_t generators_offset, const uint8_t* b_vector, const uint8_t* product,
                    const uint8_t* a_commit160, const uint8_t* l_vector, const uint8_t* r_vector,
                    const uint8_t* ap_value) {
    stream_t s = ctx.s;
    const unsigned k = ceil_log2(n);
    const uint64_t np = 1ull << k;
    Transcript tr(transcript203);
    init_transcript(tr, n);
    std::vector<Sc> x(k);
    for (unsigned j = 0; j < k; ++j)
      x[j] = round_challenge(tr, l_vector + 32 * j, r_vector + 32 * j);
    // exponents: [product', g_0 .. g_{np-1}, -x_j^2 ..., -x_j^-2 ...]
    // (prfip::compute_verification_exponents, verification_computation.cc:86-127)
    const uint64_t num = 1 + np + 2 * k;
    std::vector<uint8_t> e(32 * num);
    const Sc ap = sc_load(ap_value);
    if (n == 1) {
      sc_store(&e[0], sc_mul(sc_load(b_vector), ap));
      sc_store(&e[32], ap);
    } else {
      Sc allinv = sc_one();
      std::vector<Sc> xsq(k);
      for (unsigned j = 0; j < k; ++j) {
        Sc xi = sc_inv(x[j]);
        allinv = sc_mul(allinv, xi);
        xsq[j] = sc_mul(x[j], x[j]);
        sc_store(&e[32 * (1 + np + j)], sc_neg(xsq[j]));
        sc_store(&e[32 * (1 + np + k + j)], sc_neg(sc_mul(xi, xi)));
      }
      // g_i = ap * prod_j x_j^(+-1): bit t of i (t = 0 least significant) selects x_{k-1-t}
      std::vector<Sc> g(np);
      g[0] = sc_mul(allinv, ap);
      uint64_t filled = 1;
      for (unsigned t = 0; t < k; ++t) {
        const Sc m = sc_to_mont(xsq[k - 1 - t]);
        for (uint64_t i = 0; i < filled; ++i)
          g[filled + i] = sc_mul_mont(m, g[i]);
        filled *= 2;
      }
      Sc prod = sc_zero();
      for (uint64_t i = 0; i < n; ++i)
        prod = sc_muladd(g[i], sc_load(b_vector + 32 * i), prod);
      sc_store(&e[0], prod);
      for (uint64_t i = 0; i < np; ++i)
        sc_store(&e[32 * (1 + i)], g[i]);
    }
    // generators: [Q, G_0 .. G_{np-1}, L_j ..., R_j ...]
    DevBuf<C::Gen> gstore(np + 1, s);
    const C::Gen* G0 = generators(ctx, gstore, generators_offset, np + 1);
    DevBuf<C::Gen> gens(num, s);
    copy_d2d(gens.p, G0 + np, sizeof(C::Gen), s);
    copy_d2d(gens.p + 1, G0, np * sizeof(C::Gen), s);
    DevBuf<u32> bad(1, s);
    dev_zero(bad.p, sizeof(u32), s);
    if (k) {
      DevBuf<unsigned char> lr(64 * k, s);
      copy_h2d(lr.p, l_vector, 32 * k, s);
      copy_h2d(lr.p + 32 * k, r_vector, 32 * k, s);
      launch(DecodeToGenBody{lr.p, gens.p + 1 + np, bad.p}, 2 * k, s);
      stream_sync(s);  // lr is freed (stream-ordered) after the kernel
    }
    uint8_t expected[32], commit[32];
    msm_compressed(ctx, expected, gens.p, num, &e[0], nullptr, nullptr);
    // commit = product * Q + a_commit
    DevBuf<unsigned char> araw(C::kAbiGenBytes, s);
    DevBuf<C::Gen> pair(2, s);
    copy_h2d(araw.p, a_commit160, C::kAbiGenBytes, s);
    copy_d2d(pair.p, G0 + np, sizeof(C::Gen), s);
    launch(IngestBody<C, false>{araw.p, pair.p + 1}, 1, s);
    uint8_t sc2[64];
    std::memcpy(sc2, product, 32);
    sc_store(sc2 + 32, sc_one());
    msm_compressed(ctx, commit, pair.p, 2, sc2, nullptr, nullptr);
    u32 nbad = 0;
    copy_d2h(&nbad, bad.p, sizeof(u32), s);
    stream_sync(s);
    return (nbad == 0 && std::memcmp(expected, commit, 32) == 0) ? 1 : 0;
  }
};

}  // namespace b200
