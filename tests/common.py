"""Shared input builders for the parity tests (seeded, deterministic)."""
import numpy as np

# the reference's end-to-end golden (rust/tests/src/main.rs:21-49): three u32 columns over the
# built-in generators at offset 0 and their ristretto255 commitments
GOLDEN_COLUMNS = [[2000, 7500, 5000, 1500], [5000, 0, 400000, 10], [7000, 7500, 405000, 1510]]
GOLDEN_COMMITMENTS = [
    [4, 105, 58, 131, 59, 69, 150, 106, 120, 137, 32, 225, 175, 244, 82, 115,
     216, 180, 206, 150, 21, 250, 240, 98, 251, 192, 146, 244, 54, 169, 199, 97],
    [2, 254, 178, 195, 198, 238, 44, 156, 24, 29, 88, 196, 37, 63, 157, 50,
     236, 159, 61, 49, 153, 181, 79, 126, 55, 188, 67, 1, 228, 248, 72, 51],
    [30, 237, 163, 234, 252, 111, 45, 133, 235, 227, 21, 117, 229, 188, 88, 149,
     240, 109, 205, 90, 6, 130, 199, 152, 5, 221, 57, 231, 168, 9, 141, 122],
]
CMP = {0: 32, 1: 48, 2: 65, 3: 65}  # bytes of a commitment that are specified (no struct padding)


def golden_columns():
    return [(np.array(c, dtype="<u4").view(np.uint8).reshape(4, 4), 0) for c in GOLDEN_COLUMNS]


def random_columns(rng, n, shapes):
    """shapes: list of (length delta, element_nbytes, is_signed)."""
    cols = []
    for delta, nbytes, signed in shapes:
        m = max(0, n + delta)
        cols.append((rng.integers(0, 256, (m, nbytes), dtype=np.uint8), signed))
    return cols


def generators_for(port, curve, n, seed=3):
    """(commit-API generators, projective ABI generators) for a curve."""
    if curve == 0:
        g = port.ristretto_generators(n, seed)
        return g, g
    p2, af = port.test_points(curve, n, seed)
    return af, p2


def same(curve, a, b):
    k = CMP[curve]
    return np.array_equal(np.asarray(a)[:, :k], np.asarray(b)[:, :k])


# edge-case matrix modelled on mtxtst::exercise_multiexponentiation_fn
# (sxt/multiexp/test/multiexponentiation.cc:42-451)
def edge_case_columns():
    def u(vals, nbytes):
        return np.array([[(v >> (8 * k)) & 0xFF for k in range(nbytes)] for v in vals],
                        dtype=np.uint8).reshape(len(vals), nbytes)

    def s(vals, nbytes):
        return u([v & ((1 << (8 * nbytes)) - 1) for v in vals], nbytes)

    return [
        (u([0], 1), 0), (u([1], 1), 0), (u([2], 1), 0), (u([3], 1), 0),
        (u([0xFFFFFFFFFFFFFFFF], 8), 0),
        (u([1, 2, 3], 4), 0), (u([0, 0, 0], 4), 0),
        (u([1, 0, 255, 256, 65535], 3), 0),
        (s([-1], 1), 1), (s([-1, 1, -128, 127], 1), 1), (s([-(1 << 63), (1 << 63) - 1], 8), 1),
        (s([-(1 << 127), (1 << 127) - 1, -1, 0], 16), 1),
        (u([(1 << 256) - 1, (1 << 255), (1 << 252) + 27742317777372353535851937790883648493], 32), 0),
        (u([], 4), 0),
        (u([5] * 40, 2), 0),  # one heavily loaded bucket
    ]
