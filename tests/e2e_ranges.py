"""e2e timing (pinned host buffers through the C ABI) vs number of generator ranges."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blitzar_b200 as bb
torch.cuda.set_device(0)
bb.sxt_init()
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.empty((n, 160), dtype=torch.uint8).pin_memory(); g.numpy()[:] = bb.get_generators(n, 0)
s = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
rng = np.random.default_rng(1); s.numpy()[:] = rng.integers(0, 256, (n, 32), dtype=np.uint8); s.numpy()[:, 31] &= 0x0f
ncols = int(os.environ.get("COLS", "1"))
cols = [(s.numpy(), 0)] * ncols
ref = None
for ranges in sys.argv[2:] or ["1", "2", "4", "8"]:
    os.environ["BLITZAR_B200_RANGES"] = ranges
    for _ in range(3):
        out = bb.compute_pedersen_commitments(0, cols, g.numpy())
    t = time.perf_counter()
    for _ in range(10):
        out = bb.compute_pedersen_commitments(0, cols, g.numpy())
    dt = (time.perf_counter() - t) / 10
    if ref is None: ref = out.tobytes()
    print(f"n=2^{n.bit_length()-1} cols={ncols} ranges={ranges}: {dt*1e3:.3f} ms  {ncols*n/dt:.3e} terms/s same={out.tobytes()==ref}", flush=True)
