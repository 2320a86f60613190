"""BASELINE config C4 in ONE C-ABI call: 64 columns x 2^20 ristretto terms over explicit generators,
columns split over the GPUs of this process (BLITZAR_B200_DEVICES=k). Pinned host buffers.
Run: BLITZAR_B200_DEVICES=8 python tests/c4_full.py [ncols] [log2 n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200.api as bb  # noqa: E402

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
bb.sxt_init()
g = torch.empty((n, 160), dtype=torch.uint8).pin_memory()
g.numpy()[:] = bb.get_generators(n, 0)
host = torch.empty((ncols, n, 32), dtype=torch.uint8).pin_memory()
for c in range(ncols):
    rng = np.random.default_rng(c)
    host[c].numpy()[:] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    host[c].numpy()[:, 31] &= 0x0f
cols = [(host[c].numpy(), 0) for c in range(ncols)]
ref, best = None, 1e9
for it in range(5):
    t = time.perf_counter()
    out = bb.compute_pedersen_commitments(0, cols, g.numpy())
    dt = time.perf_counter() - t
    if it:
        best = min(best, dt)
    if ref is None:
        ref = out.copy()
    assert np.array_equal(ref, out)
# spot-check two columns against single-column calls on the primary device
for c in (0, ncols - 1):
    assert np.array_equal(bb.compute_pedersen_commitments(0, [cols[c]], g.numpy())[0], ref[c])
print(f"C4: devices={os.environ.get('BLITZAR_B200_DEVICES', '1')} {ncols} columns x 2^{n.bit_length() - 1}: "
      f"{best * 1e3:.2f} ms  {ncols * n / best:.3e} terms/s (whole call, host in -> host out)", flush=True)
