// Instantiates every kernel of the MSM engine for Ed25519 (one translation unit per curve so the
// four curves compile in parallel) and exports them through the curve's vtable.
#include "engine.cuh"
namespace b200 {
B200_DEFINE_CURVE_VTABLE(kVTableEd25519, Ed25519);
void launch_builtin_generators(const EngineCtx& ctx, void* gens, uint64_t first, uint64_t n) {
  launch(BuiltinGeneratorBody{(Ed25519::Gen*)gens, first}, n, ctx.s);
}
}  // namespace b200
