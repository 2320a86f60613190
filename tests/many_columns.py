"""Short-n / many-output shapes (SURVEY §8 a7 bucket_method2, N3 packed fixed MSM: Proof-of-SQL's call
pattern — 256 <= n <= 2^16 rows, 64 .. 1024 narrow outputs). Host-to-host through the C ABI.
    python tests/many_columns.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb  # noqa: E402
from oracle import port  # noqa: E402

bb.sxt_init(num_precomputed_generators=1 << 16)
port.build()
rng = np.random.default_rng(0)


def best_of(fn, iters=3):
    best = 1e9
    out = None
    for _ in range(iters):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


print("| call | rows n | outputs | widths | ms | terms/s |")
print("|---|---|---|---|---|---|")
for n, m in ((256, 64), (1024, 256), (4096, 1024), (16384, 1024), (65536, 64)):
    # commitments API: m columns of mixed narrow widths over the built-in generators
    widths = [(1, 0), (2, 1), (4, 0), (8, 1), (16, 0), (32, 0), (1, 0), (8, 0)]
    cols = [(rng.integers(0, 256, (n, widths[j % 8][0]), dtype=np.uint8), widths[j % 8][1]) for j in range(m)]
    ms, out = best_of(lambda: bb.compute_pedersen_commitments(0, cols, None, 0))
    if n * m <= 1 << 18:
        assert np.array_equal(out, port.commit(0, cols, None, 0))
    print(f"| sxt_curve25519_compute_pedersen_commitments | {n} | {m} | 1..32 B mixed, signed | {ms:.2f} | {n * m / ms * 1e3:.3e} |", flush=True)
for curve, name in ((0, "ristretto255"), (2, "bn254")):
    for n, m in ((1024, 256), (4096, 1024), (65536, 256)):
        gens = bb.synthetic_generators(curve, n, 0, projective=True)
        h = bb.MultiexpHandle(curve, gens)
        bt = [(1, 8, 16, 32, 64, 5, 12, 64)[j % 8] for j in range(m)]
        row = (sum(bt) + 7) // 8
        psc = rng.integers(0, 256, (n, row), dtype=np.uint8)
        ms, res = best_of(lambda: h.fixed_packed_multiexponentiation(bt, n, psc))
        if n * m <= 1 << 18:
            want = port.fixed_msm(curve, gens, m, n, psc, output_bit_table=bt)
            assert np.array_equal(port.normalize(curve, res)[:, :32], port.normalize(curve, want)[:, :32])
        print(f"| sxt_fixed_packed_multiexponentiation ({name}) | {n} | {m} | 1..64 bits packed | {ms:.2f} | {n * m / ms * 1e3:.3e} |", flush=True)
        h.free()
