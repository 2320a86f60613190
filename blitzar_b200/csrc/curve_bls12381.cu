// Instantiates every kernel of the MSM engine for Bls12381G1 (one translation unit per curve so the
// four curves compile in parallel).
#include "engine.cuh"
namespace b200 {
template struct CurveOps<Bls12381G1>;
}  // namespace b200
