// Instantiates every kernel of the MSM engine for Ed25519 (one translation unit per curve so the
// four curves compile in parallel).
#include "engine.cuh"
namespace b200 {
template struct CurveOps<Ed25519>;
void launch_builtin_generators(const EngineCtx& ctx, Ed25519::Gen* gens, uint64_t first,
                               uint64_t n) {
  launch(BuiltinGeneratorBody{gens, first}, n, ctx.s);
}
}  // namespace b200
