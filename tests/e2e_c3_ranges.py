"""C3 (bls12-381, n = 2^22) whole C-ABI call from PINNED host memory for several upload-piece counts
(BLITZAR_B200_RANGES), beside the device-resident time.   python tests/e2e_c3_ranges.py [curve] [log2 n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb  # noqa: E402

curve = int(sys.argv[1]) if len(sys.argv) > 1 else 1
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
bb.sxt_init()
n = 1 << logn
stride = {1: 104, 2: 72, 3: 72}[curve]
buf = bb.DeviceBuffer(n * stride)
bb.synthetic_generators_device(curve, buf.ptr, n, 0, False)
g = torch.empty((n, stride), dtype=torch.uint8).pin_memory()
g.numpy()[:] = buf.to_host((n, stride))
rng = np.random.default_rng(1)
s = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
s.numpy()[:] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
s.numpy()[:, 31] &= 0x3F
ref = None
for ranges in (os.environ.get("PIECES", "1,2,3,4,6,8").split(",") + [None]):
    if ranges:
        os.environ["BLITZAR_B200_RANGES"] = ranges
    else:
        os.environ.pop("BLITZAR_B200_RANGES", None)
    best = 1e9
    for it in range(4):
        t = time.perf_counter()
        out = bb.compute_pedersen_commitments(curve, [(s.numpy(), 0)], g.numpy())
        dt = time.perf_counter() - t
        if it:
            best = min(best, dt)
    ref = out if ref is None else ref
    print(f"curve {curve} n=2^{logn} pieces={ranges or 'default'}: {best * 1e3:.2f} ms  same={np.array_equal(out, ref)}", flush=True)
