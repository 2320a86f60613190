// Instantiates every kernel of the MSM engine for GrumpkinG (one translation unit per curve so the
// four curves compile in parallel).
#include "engine.cuh"
namespace b200 {
template struct CurveOps<GrumpkinG>;
}  // namespace b200
