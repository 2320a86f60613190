"""TEST INFRASTRUCTURE — Python side of the CPU emulation harness (tests/emul/emul.cpp).

Runs the product's kernel BODIES as serial host loops; used only by the `not gpu` tests to check
the pipeline logic without a GPU. The product library never links or calls any of this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
SRC = os.path.join(_HERE, "emul.cpp")
LIB_PATH = os.path.join(_HERE, "libb200_emul.so")
SIZES = {0: (160, 160, 32), 1: (144, 104, 48), 2: (96, 72, 72), 3: (96, 72, 72)}


class SequenceDescriptor(C.Structure):
    _fields_ = [("element_nbytes", C.c_uint8), ("n", C.c_uint64), ("data", C.c_void_p),
                ("is_signed", C.c_int)]


_lib = None


def build():
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from blitzar_b200 import build as b
    return b.build_emul()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emul_point_bytes.restype = C.c_uint
    return _lib


def _desc(columns):
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


def set_tuning(window_bits=0, chunk1=0, chunkn=0):
    lib().emul_set_tuning(C.c_uint(window_bits), C.c_uint(chunk1), C.c_uint(chunkn))


def set_range_entries(v=0):
    lib().emul_set_range_entries(C.c_ulonglong(v))


def set_group_entries(v=0):
    lib().emul_set_group_entries(C.c_ulonglong(v))


def set_table(window_bits=0, policy=0):
    """fixed-base table of emul_fixed handles; policy 1 = always use it, 2 = never, 0 = cost model"""
    lib().emul_set_table(C.c_uint(window_bits), C.c_uint(policy))


def set_builtin(num_precomputed=0, window_bits=0):
    lib().emul_set_builtin(C.c_uint64(num_precomputed), C.c_uint(window_bits))


def set_pairs(levels=-1, batch=0):
    """batch-affine pair levels of the Weierstrass accumulation (-1 = automatic)"""
    lib().emul_set_pairs(C.c_int(levels), C.c_uint(batch))


def set_ranges(num_ranges=1):
    lib().emul_set_ranges(C.c_uint(num_ranges))


def commit(curve_id, columns, generators=None, offset=0):
    desc, keep = _desc(columns)
    out = np.zeros((len(columns), SIZES[curve_id][2]), dtype=np.uint8)
    gp = C.c_void_p(generators.ctypes.data) if generators is not None else C.c_void_p(None)
    lib().emul_commit(C.c_uint(curve_id), C.c_void_p(out.ctypes.data), C.c_uint32(len(columns)),
                      desc, gp, C.c_uint64(offset))
    return out


def commit_partial(curve_id, columns, generators=None, offset=0):
    desc, keep = _desc(columns)
    pb = lib().emul_point_bytes(C.c_uint(curve_id))
    out = np.zeros((len(columns), pb), dtype=np.uint8)
    gp = C.c_void_p(generators.ctypes.data) if generators is not None else C.c_void_p(None)
    lib().emul_commit_partial(C.c_uint(curve_id), C.c_void_p(out.ctypes.data),
                              C.c_uint32(len(columns)), desc, gp, C.c_uint64(offset))
    return out


def combine_partials(curve_id, partials, num_parts, count):
    out = np.zeros((count, SIZES[curve_id][2]), dtype=np.uint8)
    partials = np.ascontiguousarray(partials)
    lib().emul_combine_partials(C.c_uint(curve_id), C.c_void_p(out.ctypes.data),
                                C.c_void_p(partials.ctypes.data), C.c_uint32(num_parts),
                                C.c_uint32(count))
    return out


def point_bytes(curve_id):
    return int(lib().emul_point_bytes(C.c_uint(curve_id)))


def get_generators(n, offset=0):
    out = np.zeros((n, 160), dtype=np.uint8)
    lib().emul_get_generators(C.c_void_p(out.ctypes.data), C.c_uint64(n), C.c_uint64(offset))
    return out


def fixed_msm(curve_id, generators_p, num_outputs, n, scalars, element_num_bytes=0,
              output_bit_table=None, output_lengths=None):
    res = np.zeros((num_outputs, SIZES[curve_id][0]), dtype=np.uint8)
    mode = 0 if output_bit_table is None else (1 if output_lengths is None else 2)
    bt = (C.c_uint * num_outputs)(*output_bit_table) if output_bit_table is not None else None
    ol = (C.c_uint * num_outputs)(*output_lengths) if output_lengths is not None else None
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    scalars = np.concatenate([scalars.reshape(-1), np.zeros(64, dtype=np.uint8)])
    generators_p = np.ascontiguousarray(generators_p)
    lib().emul_fixed(C.c_uint(curve_id), C.c_void_p(res.ctypes.data),
                     C.c_void_p(generators_p.ctypes.data), C.c_uint(generators_p.shape[0]),
                     C.c_int(mode), C.c_uint(element_num_bytes), bt, ol, C.c_uint(num_outputs),
                     C.c_uint(n), C.c_void_p(scalars.ctypes.data))
    return res


def check_mul(field_id, iters=2000, seed=1):
    return int(lib().emul_check_mul(C.c_uint(field_id), C.c_uint(iters), C.c_uint(seed)))


def prove_inner_product(transcript, a, b, generators_offset=0):
    """Emulated sxt_curve25519_prove_inner_product; returns (l_vector, r_vector, ap_value)."""
    n = a.shape[0]
    rounds = max(0, (n - 1).bit_length())
    lv = np.zeros((max(rounds, 1), 32), dtype=np.uint8)
    rv = np.zeros((max(rounds, 1), 32), dtype=np.uint8)
    ap = np.zeros(32, dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    lib().emul_prove_inner_product(C.c_void_p(lv.ctypes.data), C.c_void_p(rv.ctypes.data),
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
    lib().emul_verify_inner_product.restype = C.c_int
    return int(lib().emul_verify_inner_product(
        C.c_void_p(transcript.ctypes.data), C.c_uint64(n), C.c_uint64(generators_offset),
        C.c_void_p(b.ctypes.data), C.c_void_p(np.ascontiguousarray(product).ctypes.data),
        C.c_void_p(np.ascontiguousarray(a_commit).ctypes.data), C.c_void_p(lv.ctypes.data),
        C.c_void_p(rv.ctypes.data), C.c_void_p(np.ascontiguousarray(ap_value).ctypes.data)))


def synth_generators(curve_id, n, first=0, projective=False):
    stride = SIZES[curve_id][0 if (projective or curve_id == 0) else 1]
    out = np.zeros((n, stride), dtype=np.uint8)
    lib().emul_synth_generators(C.c_uint(curve_id), C.c_void_p(out.ctypes.data), C.c_uint64(n),
                                C.c_uint64(first), C.c_int(1 if projective else 0))
    return out


def generators_from_reference_table(curve_id, path):
    """Generators (projective ABI structs) recovered from a reference-format handle file."""
    raw = np.fromfile(path, dtype=np.uint8)
    w = int(raw[:4].view("<u4")[0])
    esz = {0: 120, 1: 96, 2: 64, 3: 64}[curve_id]
    groups = (raw.size - 4) // (esz << w)
    n = groups * w
    table = np.ascontiguousarray(raw[4:])
    out = np.zeros((n, SIZES[curve_id][0]), dtype=np.uint8)
    lib().emul_ingest_compact(C.c_uint(curve_id), C.c_void_p(table.ctypes.data), C.c_uint(w),
                              C.c_uint64(n), C.c_void_p(out.ctypes.data))
    return out


def set_scatter_window_major(on=0):
    lib().emul_set_scatter_window_major(C.c_uint(on))


def set_uniform_add(on=0):
    lib().emul_set_uniform_add(C.c_uint(on))


def set_range_skew(skew=0):
    lib().emul_set_range_skew(C.c_int(skew))
