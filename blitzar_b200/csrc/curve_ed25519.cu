// Instantiates every kernel of the MSM engine for Ed25519 (one translation unit per curve so the
// four curves compile in parallel) and exports them through the curve's vtable.
#include "engine.cuh"
#include "ipa.cuh"
namespace b200 {
B200_DEFINE_CURVE_VTABLE(kVTableEd25519, Ed25519);
void launch_builtin_generators(const EngineCtx& ctx, void* gens, uint64_t first, uint64_t n) {
  launch(BuiltinGeneratorBody{(Ed25519::Gen*)gens, first}, n, ctx.s);
}
unsigned selftest_lane_arithmetic(const EngineCtx& ctx, unsigned warps, unsigned seed) {
#ifdef B200_LANE_TAIL
  DevBuf<u32> bad(1, ctx.s);
  dev_zero(bad.p, sizeof(u32), ctx.s);
  launch(lane10::SelfTestBody{seed, bad.p}, (u64)warps * 32, ctx.s);
  u32 host = 0;
  copy_d2h(&host, bad.p, sizeof(u32), ctx.s);
  stream_sync(ctx.s);
  return host;
#else
  (void)ctx;
  (void)warps;
  (void)seed;
  return 0;
#endif
}
void ipa_prove(const EngineCtx& ctx, uint8_t* l_vector, uint8_t* r_vector, uint8_t* ap_value,
               uint8_t* transcript203, uint64_t n, uint64_t generators_offset,
               const uint8_t* a_vector, const uint8_t* b_vector) {
  Ipa::prove(ctx, l_vector, r_vector, ap_value, transcript203, n, generators_offset, a_vector,
             b_vector);
}
int ipa_verify(const EngineCtx& ctx, uint8_t* transcript203, uint64_t n,
               uint64_t generators_offset, const uint8_t* b_vector, const uint8_t* product,
               const uint8_t* a_commit160, const uint8_t* l_vector, const uint8_t* r_vector,
               const uint8_t* ap_value) {
  return Ipa::verify(ctx, transcript203, n, generators_offset, b_vector, product, a_commit160,
                     l_vector, r_vector, ap_value);
}
}  // namespace b200
