"""Generates the committed golden fixtures from the REFERENCE's own CPU implementation
(oracle/_ref, built from /root/reference by oracle/ref_build/Makefile). Run in the build container:

    python tests/golden/make_golden.py

Inputs are seeded; generators for the Weierstrass curves come from the reference's
fast_random_number_generator{i+1,i+2} -> generate_random_element scheme
(cbindings/pedersen.t.cc:81-123)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260922)
    n = 96
    for curve in range(4):
        if curve == 0:
            gens = refcpu.ristretto_generators(n, 7)
            gens_p = gens
        else:
            gens_p, gens = refcpu.random_elements(curve, n, first=5)
        cols = [(rng.integers(0, 256, (n, 32), dtype=np.uint8), 0),
                (rng.integers(0, 256, (n - 5, 16), dtype=np.uint8), 1),
                (rng.integers(0, 256, (n, 1), dtype=np.uint8), 0),
                (rng.integers(0, 256, (1, 8), dtype=np.uint8), 1),
                (np.zeros((0, 4), dtype=np.uint8), 0)]
        out = refcpu.commit(curve, cols, gens)
        np.savez_compressed(os.path.join(HERE, f"commit_curve{curve}.npz"), generators=gens,
                            commitments=out, signed=np.array([c[1] for c in cols]),
                            **{f"col{j}": c[0] for j, c in enumerate(cols)})
        m, outs, nb = 24, 3, 4
        sc = rng.integers(0, 256, (m, outs * nb), dtype=np.uint8)
        res = refcpu.fixed_msm(curve, gens_p[:m], outs, m, sc, element_num_bytes=nb)
        bt = [3, 11, 1, 9]
        row = (sum(bt) + 7) // 8
        psc = rng.integers(0, 256, (m, row), dtype=np.uint8)
        pres = refcpu.fixed_msm(curve, gens_p[:m], len(bt), m, psc, output_bit_table=bt)
        np.savez_compressed(os.path.join(HERE, f"fixed_curve{curve}.npz"), generators_p=gens_p[:m],
                            scalars=sc, num_outputs=outs, n=m, element_num_bytes=nb,
                            normalized=refcpu.normalize(curve, res), bit_table=np.array(bt),
                            packed_scalars=psc, packed_normalized=refcpu.normalize(curve, pres))
    g = refcpu.ristretto_generators(16, 1000)
    np.savez_compressed(os.path.join(HERE, "ristretto_generators.npz"), n=16, offset=1000,
                        compressed=refcpu.normalize(0, g))


if __name__ == "__main__":
    main()
