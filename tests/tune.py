"""Tuning sweep (run on the GPU box): device-resident MSM time vs engine parameters."""
import sys, os, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bb.sxt_init()
n = 1 << logn
rng = np.random.default_rng(0)
if curve == 0:
    gens = bb.get_generators(n, 0)
else:
    gens = bb.synthetic_generators(curve, n, 0, projective=False)  # distinct points
s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
s[:, 31] &= 0x0f
dg = bb.DeviceBuffer(host=gens); ds = bb.DeviceBuffer(host=s); do = bb.DeviceBuffer(256)

def run(iters=5):
    best = 1e9
    for it in range(iters):
        e0, e1 = bb.Event(), bb.Event()
        e0.record()
        bb.commit_device(curve, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr)
        e1.record()
        best = min(best, e0.elapsed_ms(e1))
    return best

ref = None
grid = [(c, k1, kn, g1, gn) for c in (0, 15) for k1 in (0, 48, 96) for kn in (8, 4, 16)
        for g1, gn in ((16, 4), (8, 4), (32, 4), (16, 8), (8, 8), (8, 2), (4, 4), (32, 8))]
if len(sys.argv) > 3:
    grid = [tuple(int(x) for x in a.split(',')) for a in sys.argv[3:]]
for c, k1, kn, g1, gn in grid:
    bb.set_tuning(c, k1, kn); bb.set_reduce_groups(g1, gn)
    bb.profile_accumulate(True); bb.profile_read()
    ms = run()
    acc, cnt = bb.profile_read()
    out = do.to_host()[:32].tobytes()
    if ref is None: ref = out
    print(f"c={c} k1={k1} kn={kn} g1={g1} gn={gn}: {ms:.3f} ms  acc={acc/max(cnt,1):.3f}  same={out==ref}", flush=True)
