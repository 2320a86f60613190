"""Generates tests/golden/ref_table_curve{c}_w3.bin: handle files in the REFERENCE's own format
([u32 window_width][partition table of compact elements], in_memory_partition_table_accessor.h:98-105)
written by the reference's code (oracle/_ref) for the first 7 generators of fixed_curve{c}.npz.
Run in the build container:  python tests/golden/make_table_files.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    for curve in range(4):
        g = np.load(os.path.join(HERE, f"fixed_curve{curve}.npz"))["generators_p"][:7]
        refcpu.write_partition_table(curve, os.path.join(HERE, f"ref_table_curve{curve}_w3.bin"), g, 3)
