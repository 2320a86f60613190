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

// ---- scalar work of the prover on the device (FSc25: Montgomery arithmetic mod l on 8 x u32 limbs;
// a 32-byte little-endian scalar IS an FSc25 element in plain form). Replaces the reference's
// scalar_fold_kernel / inner-product kernels (sxt/proof/inner_product/gpu_driver.cc:104-221,
// sxt/scalar25/operation/inner_product.cc). mul(x R, y) = x y keeps vectors in plain form.
typedef FSc25 FS;
typedef FS::E ScE;
B200_HD ScE sc_r2() {
  ScE r;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    r.l[i] = Sc25Params::r2(i);
  return r;
}
// any 256-bit value -> canonical residue (s25o::reduce32)
struct IpaReduceBody {
  static constexpr int kBlock = 128;
  ScE* v;
  B200_HD void operator()(u64 i) const {
    ScE t;
    FS::mul(t, v[i], sc_r2());
    FS::from_mont(v[i], t);
  }
};
// per-thread partial sums of <a_lo, b_hi> and <a_hi, b_lo> (each product carries a factor R^-1)
struct IpaDotBody {
  static constexpr int kBlock = 64;
  const ScE* a;
  const ScE* b;
  u64 mid, nl, nr;  // products in the left / right sum
  u32 K;
  ScE* partial;  // 2 per thread
  B200_HD void operator()(u64 t) const {
    ScE cl = FS::zero(), cr = FS::zero(), p;
    const u64 b0 = t * K, e0 = b0 + K;
    for (u64 i = b0; i < e0; ++i) {
      if (i < nl) {
        FS::mul(p, a[i], b[mid + i]);
        FS::add(cl, cl, p);
      }
      if (i < nr) {
        FS::mul(p, a[mid + i], b[i]);
        FS::add(cr, cr, p);
      }
    }
    partial[2 * t] = cl;
    partial[2 * t + 1] = cr;
  }
};
// sums the partials (thread 0: left, thread 1: right), removes the R^-1 and writes the scalar of Q
struct IpaDotFinalBody {
  static constexpr int kBlock = 32;
  const ScE* partial;
  u64 T;
  ScE* out_l;
  ScE* out_r;
  B200_HD void operator()(u64 side) const {
    ScE acc = FS::zero();
    for (u64 t = 0; t < T; ++t)
      FS::add(acc, acc, partial[2 * t + side]);
    FS::mul(acc, acc, sc_r2());
    *(side ? out_r : out_l) = acc;
  }
};
// out[i] = m_lo * v[i] + m_hi * v[mid + i] (v zero-padded to 2 mid; prfip::fold_scalars); the
// multipliers are in Montgomery form
struct IpaFoldScalarsBody {
  static constexpr int kBlock = 128;
  const ScE* v;
  u64 mid, hi;  // hi = number of elements in the upper half
  ScE m_lo, m_hi;
  ScE* out;
  B200_HD void operator()(u64 i) const {
    ScE t, u;
    FS::mul(t, m_lo, v[i]);
    if (i < hi) {
      FS::mul(u, m_hi, v[mid + i]);
      FS::add(t, t, u);
    }
    out[i] = t;
  }
};

// The two MSMs of a round as TWO COLUMNS of one engine call over the generator array
// [G_lo | G_hi | Q] (len + 1 entries): column L = [0 .. 0 | a_lo | c_L], column R = [a_hi, 0 .. | 0 .. 0
// | c_R]. Zero scalars produce no bucket entries, and the latency-bound tail of a small MSM (bucket
// reduction, 240 Horner doublings, ristretto encoding) is paid once per round instead of twice.
struct IpaRoundColumnsBody {
  static constexpr int kBlock = 128;
  const ScE* a;
  const ScE* cq;  // c_L, c_R
  u64 mid, a_hi;
  ScE* col_l;  // 2 mid + 1 entries each
  ScE* col_r;
  B200_HD void operator()(u64 i) const {
    const ScE zero = FS::zero();
    if (i < mid) {
      col_l[i] = zero;
      col_l[mid + i] = a[i];
      col_r[i] = i < a_hi ? a[mid + i] : zero;
      col_r[mid + i] = zero;
    } else {
      col_l[2 * mid] = cq[0];
      col_r[2 * mid] = cq[1];
    }
  }
};

// verifier exponents g_i = ap * prod_j x_j^(+-1) (prfip::compute_verification_exponents,
// sxt/proof/inner_product/verification_computation.cc:86-127): g_0 = ap * prod_j x_j^-1 and every set
// bit t of i multiplies by x_(k-1-t)^2. One thread per i, at most k products; the multipliers are in
// Montgomery form, so the values stay in plain form.
struct IpaVerifyExponentsBody {
  static constexpr int kBlock = 128;
  ScE g0;            // plain form
  const ScE* xsq_m;  // k squares of the challenges, Montgomery form
  u32 k;
  ScE* out;          // np entries
  B200_HD void operator()(u64 i) const {
    ScE g = g0;
    for (u32 t = 0; t < k; ++t)
      if ((i >> t) & 1u)
        FS::mul(g, xsq_m[k - 1 - t], g);
    out[i] = g;
  }
};
// per-thread partial sums of <g, b> (each product carries a factor R^-1; slot 2t + 1 stays zero so
// that IpaDotFinalBody can finish the sum)
struct IpaDot1Body {
  static constexpr int kBlock = 64;
  const ScE* a;
  const ScE* b;
  u64 n;
  u32 K;
  ScE* partial;
  B200_HD void operator()(u64 t) const {
    ScE acc = FS::zero(), p;
    const u64 b0 = t * K, e0 = b0 + K < n ? b0 + K : n;
    for (u64 i = b0; i < e0; ++i) {
      FS::mul(p, a[i], b[i]);
      FS::add(acc, acc, p);
    }
    partial[2 * t] = acc;
    partial[2 * t + 1] = FS::zero();
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
    launch_store_commit<C>(s, pt.p, enc.p, 1, ctx.opt.lane_tail != 0);
    copy_d2h(out32, enc.p, 32, s);
    stream_sync(s);
  }
  // both points of a round: out64_dev = [L | R] encodings; gens = [G (len) | Q]
  static void round_msms_device(const EngineCtx& ctx, unsigned char* out64_dev, const C::Gen* gens_q,
                                uint64_t len, const ScE* col_l, const ScE* col_r) {
    stream_t s = ctx.s;
    std::vector<ColumnDesc> cols(2);
    for (int j = 0; j < 2; ++j) {
      cols[j].base = (const unsigned char*)(j ? col_r : col_l);
      cols[j].row_stride = 32;
      cols[j].bit_offset = 0;
      cols[j].bit_width = 256;
      cols[j].n = (u32)(len + 1);
      cols[j].is_signed = 0;
      cols[j].first_window = cols[j].num_windows = 0;
    }
    DevBuf<C::Point> pt(2, s);
    Ops::run_columns(ctx, gens_q, cols, pt.p);
    launch_store_commit<C>(s, pt.p, out64_dev, 2, ctx.opt.lane_tail != 0);
  }
  static ScE to_device_mont(const Sc& x) {  // x R mod l as device limbs
    const Sc m = sc_to_mont(x);
    ScE r;
    for (int i = 0; i < 4; ++i) {
      r.l[2 * i] = (u32)m.v[i];
      r.l[2 * i + 1] = (u32)(m.v[i] >> 32);
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
    // folded generators: every buffer keeps Q in the slot after its last generator
    DevBuf<C::Gen> gwork(np / 2 + 1, s), gwork2(np / 4 + 2, s);
    const C::Gen* G = G0;  // G0[np] is Q already
    // a, b live in HBM for the whole proof (ping-pong halves); the host sees only L, R (64 bytes) and
    // the challenge of every round — one stream synchronisation per round, for the transcript
    DevBuf<ScE> abuf(np + np / 2 + 2, s), bbuf(np + np / 2 + 2, s);
    ScE* a = abuf.p;
    ScE* b = bbuf.p;
    ScE* a_next = abuf.p + np;
    ScE* b_next = bbuf.p + np;
    copy_h2d(a, a_vector, 32 * n, s);
    copy_h2d(b, b_vector, 32 * n, s);
    launch(IpaReduceBody{a}, n, s);
    launch(IpaReduceBody{b}, n, s);
    const u32 K = 64;
    DevBuf<ScE> partial(2 * ((np / 2 + K - 1) / K) + 2, s), cq(2, s), col_l(np + 1, s),
        col_r(np + 1, s);
    DevBuf<unsigned char> lr(64, s);
    uint64_t len = np, na = n, nb = n;
    for (unsigned round = 0; round < k; ++round) {
      const uint64_t mid = len / 2;
      const uint64_t a_hi = na - mid, b_hi = nb - mid;
      const uint64_t nl = std::min<uint64_t>(mid, b_hi), nr = std::min<uint64_t>(a_hi, mid);
      const uint64_t T = (mid + K - 1) / K;
      launch(IpaDotBody{a, b, mid, nl, nr, K, partial.p}, T, s);
      launch(IpaDotFinalBody{partial.p, T, cq.p, cq.p + 1}, 2, s);
      uint8_t* l_out = l_vector + 32 * round;
      uint8_t* r_out = r_vector + 32 * round;
      launch(IpaRoundColumnsBody{a, cq.p, mid, a_hi, col_l.p, col_r.p}, mid + 1, s);
      round_msms_device(ctx, lr.p, G, len, col_l.p, col_r.p);
      uint8_t lr_host[64];
      copy_d2h(lr_host, lr.p, 64, s);
      stream_sync(s);
      std::memcpy(l_out, lr_host, 32);
      std::memcpy(r_out, lr_host + 32, 32);
      Sc x = round_challenge(tr, l_out, r_out);
      Sc x_inv = sc_inv(x);
      const ScE xm = to_device_mont(x), xim = to_device_mont(x_inv);
      launch(IpaFoldScalarsBody{a, mid, a_hi, xm, xim, a_next}, mid, s);
      std::swap(a, a_next);
      na = mid;
      if (mid == 1)
        break;
      launch(IpaFoldScalarsBody{b, mid, b_hi, xim, xm, b_next}, mid, s);
      std::swap(b, b_next);
      nb = mid;
      FoldGeneratorsBody body;
      body.g = G;
      body.mid = (u32)mid;
      scalar_words(body.m_lo, x_inv);
      scalar_words(body.m_hi, x);
      C::Gen* gout = (G == gwork.p) ? gwork2.p : gwork.p;
      body.out = gout;
      launch(body, mid, s);
      copy_d2d(gout + mid, Q, sizeof(C::Gen), s);
      G = gout;
      len = mid;
    }
    copy_d2h(ap_value, a, 32, s);
    stream_sync(s);
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
    // (prfip::compute_verification_exponents, verification_computation.cc:86-127), built in HBM: the
    // host computes only the 2k + 1 values that depend on the challenges alone
    const uint64_t num = 1 + np + 2 * k;
    DevBuf<ScE> e(num + 1, s);
    const Sc ap = sc_load(ap_value);
    if (n == 1) {
      uint8_t e2[64];
      sc_store(e2, sc_mul(sc_load(b_vector), ap));
      sc_store(e2 + 32, ap);
      copy_h2d(e.p, e2, 64, s);
      stream_sync(s);
    } else {
      Sc allinv = sc_one();
      std::vector<uint8_t> tail(64 * k), xsq_m(32 * k);
      for (unsigned j = 0; j < k; ++j) {
        const Sc xi = sc_inv(x[j]);
        allinv = sc_mul(allinv, xi);
        const Sc xs = sc_mul(x[j], x[j]);
        sc_store(&tail[32 * j], sc_neg(xs));
        sc_store(&tail[32 * (k + j)], sc_neg(sc_mul(xi, xi)));
        sc_store(&xsq_m[32 * j], sc_to_mont(xs));
      }
      DevBuf<ScE> d_xsq(k, s), d_b(n, s);
      copy_h2d(d_xsq.p, xsq_m.data(), 32 * k, s);
      copy_h2d(e.p + 1 + np, tail.data(), 64 * k, s);
      copy_h2d(d_b.p, b_vector, 32 * n, s);
      launch(IpaReduceBody{d_b.p}, n, s);
      ScE g0;
      {
        const Sc g0s = sc_mul(allinv, ap);
        for (int i = 0; i < 4; ++i) {
          g0.l[2 * i] = (u32)g0s.v[i];
          g0.l[2 * i + 1] = (u32)(g0s.v[i] >> 32);
        }
      }
      launch(IpaVerifyExponentsBody{g0, d_xsq.p, k, e.p + 1}, np, s);
      const u32 K = 64;
      const uint64_t T = (n + K - 1) / K;
      DevBuf<ScE> partial(2 * T + 2, s), scratch(1, s);
      launch(IpaDot1Body{e.p + 1, d_b.p, n, K, partial.p}, T, s);
      launch(IpaDotFinalBody{partial.p, T, e.p, scratch.p}, 2, s);
      stream_sync(s);  // the host staging vectors and the DevBufs of this scope end here
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
    {
      std::vector<ColumnDesc> cols(1);
      cols[0].base = (const unsigned char*)e.p;
      cols[0].row_stride = 32;
      cols[0].bit_offset = 0;
      cols[0].bit_width = 256;
      cols[0].n = (u32)num;
      cols[0].is_signed = 0;
      cols[0].first_window = cols[0].num_windows = 0;
      DevBuf<C::Point> pt(1, s);
      DevBuf<unsigned char> enc(32, s);
      Ops::run_columns(ctx, gens.p, cols, pt.p);
      launch_store_commit<C>(s, pt.p, enc.p, 1, ctx.opt.lane_tail != 0);
      copy_d2h(expected, enc.p, 32, s);
      stream_sync(s);
    }
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
