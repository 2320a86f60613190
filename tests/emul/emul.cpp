// TEST INFRASTRUCTURE — never linked into the product.
//
// Builds the product's kernel bodies (blitzar_b200/csrc/*.cuh) as SERIAL HOST LOOPS
// (-DB200_EMULATE: launch() iterates the thread index, device memory is malloc) so that the whole
// MSM pipeline — digit recoding, counting sort, chunked accumulation cascade, bucket reduction,
// window combination, canonicalisation — can be checked against the oracle in a container without
// a GPU. The product library (api.cu) has no such path and aborts without a GPU.
//
// The per-curve translation units of the product (csrc/curve_*.cu) are compiled as host C++ with
// emul_prefix.h force-included and reached through the same type-erased vtables api.cu uses.
#include "emul_prefix.h"
#include "../../blitzar_b200/csrc/engine_api.cuh"

using namespace b200;

static const CurveVTable& vt(unsigned curve_id) {
  switch (curve_id) {
  case 0: return kVTableEd25519;
  case 1: return kVTableBls12381;
  case 2: return kVTableBn254;
  default: return kVTableGrumpkin;
  }
}

static MsmOptions g_opt;
static unsigned g_ranges = 1;
static unsigned g_table_c = 0;  // fixed-base table window of emul_fixed handles (0 = no table)
static std::vector<unsigned char> g_builtin;  // built-in generator table of emul_commit
static uint64_t g_num_builtin = 0;
static unsigned g_builtin_c = 0;
static EngineCtx make_ctx() {
  EngineCtx ctx{0, g_opt, g_builtin.empty() ? nullptr : g_builtin.data(), g_num_builtin};
  ctx.builtin_window_bits = g_builtin_c;
  ctx.builtin_windows = g_builtin_c ? 256 / g_builtin_c + 1 : (g_num_builtin ? 1 : 0);
  return ctx;
}

extern "C" {
void emul_set_ranges(unsigned num_ranges) { g_ranges = num_ranges ? num_ranges : 1; }
void emul_set_table(unsigned window_bits, unsigned policy) {
  g_table_c = window_bits;
  g_opt.table_policy = policy;
}
// sxt_config::num_precomputed_generators with a fixed-base table of the given window (0 = none)
void emul_set_builtin(uint64_t np, unsigned window_bits) {
  g_builtin.clear();
  g_num_builtin = 0;
  g_builtin_c = 0;
  if (np == 0) return;
  const unsigned windows = window_bits ? 256 / window_bits + 1 : 1;
  std::vector<unsigned char> t((size_t)np * windows * vt(0).gen_bytes);
  EngineCtx ctx{0, g_opt, nullptr, 0};
  launch_builtin_generators(ctx, t.data(), 0, np);
  vt(0).build_table(ctx, t.data(), np, window_bits, windows);
  g_builtin.swap(t);
  g_num_builtin = np;
  g_builtin_c = window_bits;
}
void emul_set_range_entries(unsigned long long v) { g_opt.max_range_entries = v ? v : (1ull << 31); }
void emul_set_group_entries(unsigned long long v) { g_opt.max_group_entries = v ? v : (1ull << 30); }
void emul_set_scatter_window_major(unsigned on) { g_opt.scatter_window_major = on; }
void emul_set_uniform_add(unsigned on) { g_opt.uniform_add = on; }
void emul_set_range_skew(int skew) { g_opt.range_skew = skew; }
void emul_set_pairs(int levels, unsigned batch) {
  g_opt.pair_levels = levels;
  g_opt.pair_batch = batch;
}
void emul_set_tuning(unsigned window_bits, unsigned chunk1, unsigned chunkn) {
  g_opt.window_bits = window_bits;
  g_opt.chunk1 = chunk1;
  g_opt.chunkn = chunkn ? chunkn : 8;
}
// same contract as b200_commit_device, with host pointers standing in for device pointers
void emul_commit(unsigned curve_id, void* out_commitments, uint32_t num,
                 const sxt_sequence_descriptor* d, const void* generators, uint64_t offset) {
  if (num == 0) return;
  EngineCtx ctx = make_ctx();
  vt(curve_id).commit_device(ctx, out_commitments, nullptr, num, d, generators, offset, g_ranges,
                             nullptr, nullptr);
}
void emul_get_generators(void* out160, uint64_t num, uint64_t offset) {
  EngineCtx ctx = make_ctx();
  std::vector<unsigned char> g((num ? num : 1) * vt(0).gen_bytes);
  launch_builtin_generators(ctx, g.data(), offset, num);
  vt(0).gens_to_projective(ctx, g.data(), out160, num);
}
// same contract as b200_synthetic_generators_device (host memory)
void emul_synth_generators(unsigned curve_id, void* out, uint64_t n, uint64_t first, int projective) {
  EngineCtx ctx = make_ctx();
  vt(curve_id).synth_generators(ctx, out, n, first, projective != 0);
}
// generators recovered from a reference partition-table image, as projective ABI structs
void emul_ingest_compact(unsigned curve_id, const void* table, unsigned window_width, uint64_t n,
                         void* out_proj) {
  EngineCtx ctx = make_ctx();
  const CurveVTable& V = vt(curve_id);
  std::vector<unsigned char> gens((size_t)(n ? n : 1) * V.gen_bytes);
  V.ingest_compact_table(ctx, table, window_width, gens.data(), n);
  V.gens_to_projective(ctx, gens.data(), out_proj, n);
}
// handle_new + fixed MSM in one call (mode as in b200_fixed_msm_device)
void emul_fixed(unsigned curve_id, void* res, const void* generators_proj, unsigned num_gens,
                int mode, unsigned element_num_bytes, const unsigned* bit_table,
                const unsigned* lengths, unsigned num_outputs, unsigned n, const uint8_t* scalars) {
  EngineCtx ctx = make_ctx();
  const CurveVTable& V = vt(curve_id);
  const unsigned windows = g_table_c ? 256 / g_table_c + 1 : 1;
  std::vector<unsigned char> gens((size_t)(num_gens ? num_gens : 1) * windows * V.gen_bytes);
  V.ingest_projective(ctx, generators_proj, gens.data(), num_gens);
  V.build_table(ctx, gens.data(), num_gens, g_table_c, windows);
  Handle h{curve_id, num_gens, gens.data()};
  h.window_bits = g_table_c;
  h.windows = windows;
  unsigned rows = n;
  if (mode == 2) {
    rows = 0;
    for (unsigned j = 0; j < num_outputs; ++j) rows = lengths[j] > rows ? lengths[j] : rows;
  }
  V.fixed_device(ctx, res, nullptr, &h, mode, element_num_bytes, bit_table, lengths, num_outputs,
                 rows, scalars);
}
}

// field-level cross-check: production multiply schedule vs the plain reference schedule
template <class F> static int check_field(unsigned iters, unsigned seed) {
  typename F::E a, b, r1, r2;
  unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto next = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (u32)(st >> 16); };
  int bad = 0;
  for (unsigned it = 0; it < iters; ++it) {
    for (int i = 0; i < F::N; ++i) { a.l[i] = next(); b.l[i] = next(); }
    if (it % 7 == 0) for (int i = 0; i < F::N; ++i) a.l[i] = 0xffffffffu;
    if (it % 11 == 0) for (int i = 0; i < F::N; ++i) b.l[i] = 0xffffffffu;
    if (it % 13 == 0) for (int i = 0; i < F::N; ++i) b.l[i] = 0;
    // Montgomery operands are residues < p: clear the top bits (p > 2^253 resp. 2^380)
    a.l[F::N - 1] &= 0x0fffffffu;
    b.l[F::N - 1] &= 0x0fffffffu;
    if (it % 17 == 0) { a = F::modulus(); limbs_sub_small<F::N>(a.l, a.l, 1); }
    if (it % 19 == 0) { b = F::modulus(); limbs_sub_small<F::N>(b.l, b.l, 1 + it % 3); }
    F::mul(r1, a, b);
    F::mul_ref(r2, a, b);
    typename F::E c1 = r1, c2 = r2;
    bool same = true;
    for (int i = 0; i < F::N; ++i) same = same && (c1.l[i] == c2.l[i]);
    if (!same) ++bad;
  }
  return bad;
}
extern "C" int emul_check_mul(unsigned field_id, unsigned iters, unsigned seed) {
  switch (field_id) {
  case 0: {
    // F25519 results are only congruent mod p; compare canonical forms
    int bad = 0;
    F25519::E a, b, r1, r2, c1, c2;
    unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto next = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (u32)(st >> 16); };
    for (unsigned it = 0; it < iters; ++it) {
      for (int i = 0; i < 8; ++i) { a.l[i] = next(); b.l[i] = next(); }
      if (it % 7 == 0) for (int i = 0; i < 8; ++i) a.l[i] = 0xffffffffu;
      if (it % 11 == 0) for (int i = 0; i < 8; ++i) b.l[i] = 0xffffffffu;
      // Karatsuba corner cases: equal / zero halves, halves in either order
      if (it % 13 == 0) for (int i = 0; i < 4; ++i) a.l[4 + i] = a.l[i];
      if (it % 17 == 0) for (int i = 0; i < 4; ++i) b.l[i] = b.l[4 + i];
      if (it % 19 == 0) for (int i = 0; i < 4; ++i) a.l[i] = 0;
      if (it % 23 == 0) for (int i = 0; i < 4; ++i) b.l[4 + i] = 0;
      if (it % 29 == 0) for (int i = 0; i < 8; ++i) a.l[i] = 0;
      F25519::mul(r1, a, b);
      F25519::mul_ref(r2, a, b);
      F25519::canonical(c1, r1);
      F25519::canonical(c2, r2);
      for (int i = 0; i < 8; ++i) if (c1.l[i] != c2.l[i]) { ++bad; break; }
      F25519::mul_lat(r1, a, b);
      F25519::canonical(c1, r1);
      for (int i = 0; i < 8; ++i) if (c1.l[i] != c2.l[i]) { ++bad; break; }
    }
    return bad;
  }
  case 1: return check_field<FBls>(iters, seed);
  case 2: return check_field<FBn>(iters, seed);
  default: return check_field<FGk>(iters, seed);
  }
}

// binary-Euclid inversion vs the Fermat power (Montgomery fields), incl. 0, 1, 2, p - 1
template <class F> static int check_invert(unsigned iters, unsigned seed) {
  unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 3;
  auto next = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (u32)(st >> 16); };
  int bad = 0;
  for (unsigned it = 0; it < iters; ++it) {
    typename F::E a, r1, r2, one = F::one(), chk;
    for (int i = 0; i < F::N; ++i) a.l[i] = next();
    a.l[F::N - 1] &= 0x0fffffffu;
    if (it == 0) a = F::zero();
    if (it == 1) a = F::one();
    if (it == 2) F::add(a, one, one);
    if (it == 3) { a = F::modulus(); limbs_sub_small<F::N>(a.l, a.l, 1); }
    if (it == 4) { a = F::zero(); a.l[0] = 1; }  // plain 1 = R^-1 in the Montgomery domain
    F::invert(r1, a);
    F::invert_eea(r2, a);
    if (!F::equal(r1, r2)) ++bad;
    if (it) { F::mul(chk, r2, a); if (!F::equal(chk, one)) ++bad; }
  }
  return bad;
}
extern "C" int emul_check_invert(unsigned field_id, unsigned iters, unsigned seed) {
  switch (field_id) {
  case 1: return check_invert<FBls>(iters, seed);
  case 2: return check_invert<FBn>(iters, seed);
  default: return check_invert<FGk>(iters, seed);
  }
}

// multi-GPU host logic support: partial accumulator points and their combination
extern "C" unsigned emul_point_bytes(unsigned curve_id) {
  return vt(curve_id).point_bytes;
}
extern "C" void emul_commit_partial(unsigned curve_id, void* out_partials, uint32_t num,
                                    const sxt_sequence_descriptor* d, const void* generators,
                                    uint64_t offset) {
  if (num == 0) return;
  EngineCtx ctx = make_ctx();
  vt(curve_id).commit_device(ctx, nullptr, out_partials, num, d, generators, offset, g_ranges,
                             nullptr, nullptr);
}
extern "C" void emul_combine_partials(unsigned curve_id, void* out_commitments, const void* partials,
                                      uint32_t num_parts, uint32_t count) {
  EngineCtx ctx = make_ctx();
  const CurveVTable& V = vt(curve_id);
  std::vector<unsigned char> sum((size_t)count * V.point_bytes);
  V.sum_parts(ctx, partials, num_parts, count, sum.data());
  V.store(ctx, sum.data(), out_commitments, count, true);
}

// inner-product argument through the emulated kernels (same contracts as the sxt_* entry points)
extern "C" void emul_prove_inner_product(uint8_t* l_vector, uint8_t* r_vector, uint8_t* ap_value,
                                         uint8_t* transcript203, uint64_t n, uint64_t offset,
                                         const uint8_t* a_vector, const uint8_t* b_vector) {
  EngineCtx ctx = make_ctx();
  ipa_prove(ctx, l_vector, r_vector, ap_value, transcript203, n, offset, a_vector, b_vector);
}
extern "C" int emul_verify_inner_product(uint8_t* transcript203, uint64_t n, uint64_t offset,
                                         const uint8_t* b_vector, const uint8_t* product,
                                         const uint8_t* a_commit160, const uint8_t* l_vector,
                                         const uint8_t* r_vector, const uint8_t* ap_value) {
  EngineCtx ctx = make_ctx();
  return ipa_verify(ctx, transcript203, n, offset, b_vector, product, a_commit160, l_vector,
                    r_vector, ap_value);
}
