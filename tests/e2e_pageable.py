"""e2e through the C ABI from PAGEABLE numpy buffers (what a typical caller passes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb
bb.sxt_init()
rng = np.random.default_rng(1)
for logn in [int(a) for a in sys.argv[1:]] or (18, 20, 22, 24):
    n = 1 << logn
    g = bb.get_generators(n, 0)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8); s[:, 31] &= 0x0f
    for _ in range(2):
        out = bb.compute_pedersen_commitments(0, [(s, 0)], g)
    t = time.perf_counter()
    for _ in range(3):
        out = bb.compute_pedersen_commitments(0, [(s, 0)], g)
    dt = (time.perf_counter() - t) / 3
    print(f"stager threads {os.environ.get('BLITZAR_B200_STAGER_THREADS', 'default')} ristretto n=2^{logn}: e2e (pageable) {dt*1e3:.2f} ms  {n/dt:.3e} terms/s  [{192*n/dt/1e9:.1f} GB/s of input]", flush=True)
