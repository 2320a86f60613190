// Instantiates every kernel of the MSM engine for GrumpkinG (one translation unit per curve so the
// four curves compile in parallel) and exports them through the curve's vtable.
#include "engine.cuh"
namespace b200 {
B200_DEFINE_CURVE_VTABLE(kVTableGrumpkin, GrumpkinG);
}  // namespace b200
