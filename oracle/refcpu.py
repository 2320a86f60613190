"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libblitzar_ref_cpu.so.

That library is the reference's own CPU MSM path (unmodified sources compiled by
oracle/ref_build/Makefile) behind the thin `ref_*` driver in oracle/ref_build/ref_driver.cc.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libblitzar_ref_cpu.so")

CURVE_RISTRETTO255, CURVE_BLS_381, CURVE_BN_254, CURVE_GRUMPKIN = 0, 1, 2, 3
# (projective, affine-as-passed-to-commit, normalised-output) byte sizes per curve
SIZES = {0: (160, 160, 32), 1: (144, 104, 48), 2: (96, 72, 72), 3: (96, 72, 72)}


class SequenceDescriptor(C.Structure):
    """Layout of sxt_sequence_descriptor (cbindings/blitzar_api.h:115-131)."""
    _fields_ = [("element_nbytes", C.c_uint8), ("n", C.c_uint64),
                ("data", C.c_void_p), ("is_signed", C.c_int)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_sizeof.restype = C.c_uint
    return _lib


def make_descriptors(columns):
    """columns: list of (np.ndarray uint8 [n, nbytes] C-contiguous, is_signed)."""
    arr = (SequenceDescriptor * max(1, len(columns)))()
    keep = []
    for i, (data, is_signed) in enumerate(columns):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        keep.append(data)
        arr[i].element_nbytes = data.shape[1]
        arr[i].n = data.shape[0]
        arr[i].data = data.ctypes.data if data.shape[0] else None
        arr[i].is_signed = int(is_signed)
    return arr, keep


def ristretto_generators(n, offset=0):
    out = np.zeros((n, 160), dtype=np.uint8)
    if n:
        lib().ref_ristretto255_get_generators(C.c_void_p(out.ctypes.data), C.c_uint64(n),
                                              C.c_uint64(offset))
    return out


def random_elements(curve_id, n, first=0, affine=True):
    psz, asz, _ = SIZES[curve_id]
    p2 = np.zeros((n, psz), dtype=np.uint8)
    af = np.zeros((n, asz), dtype=np.uint8)
    lib().ref_random_elements(C.c_uint(curve_id), C.c_void_p(p2.ctypes.data),
                              C.c_void_p(af.ctypes.data), C.c_uint64(n), C.c_uint64(first))
    return (p2, af)


def commit(curve_id, columns, generators=None, offset=0):
    """Reference cpu-backend commitments. Returns uint8 [num_columns, out_size]."""
    desc, keep = make_descriptors(columns)
    out = np.zeros((len(columns), SIZES[curve_id][2]), dtype=np.uint8)
    gp = C.c_void_p(generators.ctypes.data) if generators is not None else C.c_void_p(None)
    L = lib()
    if curve_id == 0:
        L.ref_curve25519_commit(C.c_void_p(out.ctypes.data), C.c_uint32(len(columns)), desc, gp,
                                C.c_uint64(offset))
    else:
        fn = {1: L.ref_bls12_381_g1_commit, 2: L.ref_bn254_g1_commit,
              3: L.ref_grumpkin_commit}[curve_id]
        fn(C.c_void_p(out.ctypes.data), C.c_uint32(len(columns)), desc, gp)
    return out


def normalize(curve_id, projective):
    n = projective.shape[0]
    out = np.zeros((n, SIZES[curve_id][2]), dtype=np.uint8)
    lib().ref_normalize(C.c_uint(curve_id), C.c_void_p(out.ctypes.data),
                        C.c_void_p(projective.ctypes.data), C.c_uint64(n))
    return out


def fixed_msm(curve_id, generators_p, num_outputs, n, scalars, element_num_bytes=0,
              output_bit_table=None, output_lengths=None, window_width=4):
    """Reference cpu fixed-base MSM (builds the partition table each call). Returns projective."""
    res = np.zeros((num_outputs, SIZES[curve_id][0]), dtype=np.uint8)
    mode = 0 if output_bit_table is None else (1 if output_lengths is None else 2)
    bt = (C.c_uint * num_outputs)(*output_bit_table) if output_bit_table is not None else None
    ol = (C.c_uint * num_outputs)(*output_lengths) if output_lengths is not None else None
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    lib().ref_fixed_msm(C.c_uint(curve_id), C.c_void_p(res.ctypes.data),
                        C.c_void_p(generators_p.ctypes.data), C.c_uint(generators_p.shape[0]),
                        C.c_uint(window_width), C.c_int(mode), C.c_uint(element_num_bytes), bt, ol,
                        C.c_uint(num_outputs), C.c_uint(n), C.c_void_p(scalars.ctypes.data))
    return res


def write_partition_table(curve_id, filename, generators_p, window_width):
    """The reference's handle file ([u32 window_width][partition table]) for these generators."""
    generators_p = np.ascontiguousarray(generators_p, dtype=np.uint8)
    lib().ref_write_partition_table(C.c_uint(curve_id), filename.encode(),
                                    C.c_void_p(generators_p.ctypes.data),
                                    C.c_uint(generators_p.shape[0]), C.c_uint(window_width))


# ---- inner-product argument (reference cpu backend; cbindings/inner_product_proof.cc) -------------
def transcript_new(label=b"ip-test"):
    t = np.zeros(203, dtype=np.uint8)
    lib().ref_transcript_new(C.c_void_p(t.ctypes.data), C.c_char_p(label))
    return t


def transcript_challenge(transcript, label=b"chk"):
    out = np.zeros(32, dtype=np.uint8)
    lib().ref_transcript_challenge(C.c_void_p(out.ctypes.data), C.c_void_p(transcript.ctypes.data),
                                   C.c_char_p(label))
    return out


def prove_inner_product(transcript, a, b, generators_offset=0):
    """a, b: uint8 [n, 32] scalars (reduced mod l). Returns (l_vector, r_vector, ap_value);
    `transcript` (203 bytes) is advanced in place."""
    n = a.shape[0]
    rounds = max(0, (n - 1).bit_length())
    lv = np.zeros((max(rounds, 1), 32), dtype=np.uint8)
    rv = np.zeros((max(rounds, 1), 32), dtype=np.uint8)
    ap = np.zeros(32, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    lib().ref_prove_inner_product(C.c_void_p(lv.ctypes.data), C.c_void_p(rv.ctypes.data),
                                  C.c_void_p(ap.ctypes.data), C.c_void_p(transcript.ctypes.data),
                                  C.c_uint64(n), C.c_uint64(generators_offset),
                                  C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
    return lv[:rounds], rv[:rounds], ap


def verify_inner_product(transcript, b, product, a_commit, l_vector, r_vector, ap_value,
                         generators_offset=0):
    n = b.shape[0]
    b = np.ascontiguousarray(b, dtype=np.uint8)
    lv = np.ascontiguousarray(l_vector if len(l_vector) else np.zeros((1, 32), np.uint8))
    rv = np.ascontiguousarray(r_vector if len(r_vector) else np.zeros((1, 32), np.uint8))
    lib().ref_verify_inner_product.restype = C.c_int
    return int(lib().ref_verify_inner_product(
        C.c_void_p(transcript.ctypes.data), C.c_uint64(n), C.c_uint64(generators_offset),
        C.c_void_p(b.ctypes.data), C.c_void_p(np.ascontiguousarray(product).ctypes.data),
        C.c_void_p(np.ascontiguousarray(a_commit).ctypes.data), C.c_void_p(lv.ctypes.data),
        C.c_void_p(rv.ctypes.data), C.c_void_p(np.ascontiguousarray(ap_value).ctypes.data)))
