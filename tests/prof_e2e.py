"""Profiling driver (used under ncu): C2 through the host C ABI (pinned inputs), BLITZAR_B200_RANGES
pieces; prints nothing timed — read the ncu launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blitzar_b200 as bb
bb.sxt_init()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
g = torch.empty((n, 160), dtype=torch.uint8).pin_memory(); g.numpy()[:] = bb.get_generators(n, 0)
s = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
rng = np.random.default_rng(1); s.numpy()[:] = rng.integers(0, 256, (n, 32), dtype=np.uint8); s.numpy()[:, 31] &= 0x0f
for _ in range(2):
    out = bb.compute_pedersen_commitments(0, [(s.numpy(), 0)], g.numpy())
print("MARK launches", bb.launch_count(), flush=True)
