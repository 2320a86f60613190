"""Profiling driver (used under ncu): C2 = ristretto MSM, n = 2^20, inputs resident in HBM."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
curve = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bb.sxt_init()
n = 1 << logn
rng = np.random.default_rng(0)
if curve == 0:
    gens = bb.get_generators(n, 0)
else:
    gens = bb.synthetic_generators(curve, n, 0, projective=False)  # distinct points
s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
s[:, 31] &= 0x0f
dg = bb.DeviceBuffer(host=gens)
ds = bb.DeviceBuffer(host=s)
do = bb.DeviceBuffer(256)
for it in range(iters):
    e0, e1 = bb.Event(), bb.Event()
    e0.record()
    bb.commit_device(curve, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr)
    e1.record()
    print("iter", it, "ms", e0.elapsed_ms(e1), flush=True)
