"""Timing of the in-process multi-GPU column split (BLITZAR_B200_DEVICES): 8 and 32 columns of
n = 2^20 32-byte scalars over the built-in ristretto generators, pinned host buffers.
Run: BLITZAR_B200_DEVICES=k python tests/e2e_devices.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200.api as bb  # noqa: E402

n = 1 << 20
bb.sxt_init(num_precomputed_generators=n)
for ncols in (8, 32):
    host = torch.randint(0, 256, (ncols, n, 32), dtype=torch.uint8).pin_memory()
    cols = [(host[i].numpy(), 0) for i in range(ncols)]
    ref = None
    for it in range(4):
        t = time.perf_counter()
        out = bb.compute_pedersen_commitments(0, cols)
        dt = time.perf_counter() - t
        if ref is None:
            ref = out.copy()
        assert np.array_equal(ref, out)
    print(f"devices={os.environ.get('BLITZAR_B200_DEVICES', '1')} cols={ncols} n=2^20: "
          f"{dt * 1e3:.2f} ms  {ncols * n / dt:.3e} terms/s", flush=True)
