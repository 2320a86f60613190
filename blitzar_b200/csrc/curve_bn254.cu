// Instantiates every kernel of the MSM engine for Bn254G1 (one translation unit per curve so the
// four curves compile in parallel).
#include "engine.cuh"
namespace b200 {
template struct CurveOps<Bn254G1>;
}  // namespace b200
