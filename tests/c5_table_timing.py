"""C5 per-GPU share (bn254 fixed-base MSM, n = 2^21 by default; distinct reference generators):
the handle's fixed-base table (shared bucket set, no Horner tail) against the variable-base run over
the same resident generators, device-resident and through the host call.
    python tests/c5_table_timing.py [log2 n] [curve]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 21
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bb.sxt_init()
n = 1 << logn
stride = {0: 160, 1: 144, 2: 96, 3: 96}[curve]
rng = np.random.default_rng(0)
s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
s[:, 31] &= 0x3F if curve != 0 else 0x0F
dg = bb.DeviceBuffer(n * stride)
bb.synthetic_generators_device(curve, dg.ptr, n, 0, True)
bb.synchronize()
ds = bb.DeviceBuffer(host=s)
do = bb.DeviceBuffer(512)
results = {}
for window in (None, "16", "13"):
    if window:
        os.environ["BLITZAR_B200_TABLE_WINDOW"] = window
    t = time.perf_counter()
    h = bb.MultiexpHandle(curve, device_ptr=dg.ptr, n=n)
    bb.synchronize()
    t_new = time.perf_counter() - t
    for policy in ("1", "2"):
        os.environ["BLITZAR_B200_TABLE_POLICY"] = policy
        best = 1e9
        for _ in range(4):
            e0, e1 = bb.Event(), bb.Event()
            e0.record()
            bb.fixed_msm_device(h, do.ptr, None, 32, 1, n, ds.ptr)
            e1.record()
            best = min(best, e0.elapsed_ms(e1))
        res = do.to_host()[:stride].copy()
        t0 = time.perf_counter()
        h.fixed_multiexponentiation(32, 1, n, s)
        host_ms = (time.perf_counter() - t0) * 1e3
        results[(window, policy)] = res
        print(f"curve {curve} n=2^{logn} table_window={window or 'auto'} policy={'table' if policy == '1' else 'variable-base'}: "
              f"device {best:.3f} ms ({n / best * 1e3:.3e} terms/s), host call (pageable) {host_ms:.1f} ms, "
              f"handle_new {t_new * 1e3:.0f} ms", flush=True)
    h.free()
from oracle import refcpu  # noqa: E402
norm = {k: refcpu.normalize(curve, v.reshape(1, -1)).tobytes() for k, v in results.items()}
print("all variants agree:", len(set(norm.values())) == 1)
