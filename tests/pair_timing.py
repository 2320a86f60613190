"""Batch-affine pair levels (Weierstrass accumulation) on/off: device-resident MSM time and the
level-1 accumulation stage time.   python tests/pair_timing.py [curve] [log2 n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb  # noqa: E402

curve = int(sys.argv[1]) if len(sys.argv) > 1 else 1
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
configs = sys.argv[3:] or ["0:0", "-1:0", "-1:16", "-1:64", "3:32", "4:32", "5:32", "6:32"]
bb.sxt_init()
n = 1 << logn
stride = {1: 104, 2: 72, 3: 72}[curve]
rng = np.random.default_rng(0)
s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
s[:, 31] &= 0x7F if curve == 1 else 0x3F
dg = bb.DeviceBuffer(n * stride)
bb.synthetic_generators_device(curve, dg.ptr, n, 0, False)
ds = bb.DeviceBuffer(host=s)
do = bb.DeviceBuffer(256)
res = {}
for cfg in configs:
    lv, batch = cfg.split(":")
    os.environ["BLITZAR_B200_PAIR_LEVELS"] = lv
    os.environ["BLITZAR_B200_PAIR_BATCH"] = batch
    best, acc = 1e9, 1e9
    for it in range(3):
        bb.profile_accumulate(True)
        bb.profile_read()
        e0, e1 = bb.Event(), bb.Event()
        e0.record()
        bb.commit_device(curve, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr)
        e1.record()
        ms = e0.elapsed_ms(e1)
        a, cnt = bb.profile_read()
        bb.profile_accumulate(False)
        if ms < best:
            best, acc = ms, a
    res[cfg] = do.to_host()[:48].tobytes()
    print(f"curve {curve} n=2^{logn} pair_levels={lv} batch={batch}: {best:.3f} ms ({n / best * 1e3:.3e} terms/s), "
          f"accumulation stage {acc:.3f} ms", flush=True)
print("all agree:", len(set(res.values())) == 1)
