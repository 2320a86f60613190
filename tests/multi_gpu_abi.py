"""The BASELINE multi-GPU configs as ONE C-ABI call each from ONE process (what a Rust caller gets
with BLITZAR_B200_DEVICES=k): the library shards inside —
  C2   sxt_curve25519_compute_pedersen_commitments_with_generators, 1 column, n = 2^20 (strong scaling:
       the generator range is split over the devices, partial points gathered on device 0)
  C4   64 columns x 2^20 (split by column, no exchange)
  C5   sxt_fixed_multiexponentiation over a bn254 handle, n = 2^24 (handle sharded at construction)
with parity checks: C2 / C4 against the single-device result of the same process' oracle-checked
path, C5 against the closed form over the reference generators (one reference scalar multiplication).
Run: BLITZAR_B200_DEVICES=8 python tests/multi_gpu_abi.py [c2] [c4] [c5]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200.api as bb  # noqa: E402
from oracle import refcpu  # noqa: E402
from tests import common  # noqa: E402

which = sys.argv[1:] or ["c2", "c4", "c5"]
devices = os.environ.get("BLITZAR_B200_DEVICES", "1")
bb.sxt_init()


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.uint8).pin_memory()
    t.numpy()[:] = a
    return t


def best_of(fn, iters=4):
    best, out = 1e9, None
    for it in range(iters):
        t = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t
        if it:
            best = min(best, dt)
    return best, out


if "c2" in which:
    n = 1 << 20
    g = pinned(bb.get_generators(n, 0))
    s = pinned(common.mt19937_bytes(0, n))
    dt, out = best_of(lambda: bb.compute_pedersen_commitments(0, [(s.numpy(), 0)], g.numpy()))
    want = refcpu.commit(0, [(s.numpy(), 0)], g.numpy())  # full-size reference MSM (~20 s)
    print(f"C2 devices={devices}: {dt * 1e3:.2f} ms  {n / dt:.3e} terms/s, equals the reference cpu backend: "
          f"{np.array_equal(out, want)}", flush=True)
if "c4" in which:
    n, ncols = 1 << 20, 64
    g = pinned(bb.get_generators(n, 0))
    host = torch.empty((ncols, n, 32), dtype=torch.uint8).pin_memory()
    for c in range(ncols):
        host[c].numpy()[:] = common.mt19937_bytes(c, n)
    cols = [(host[c].numpy(), 0) for c in range(ncols)]
    dt, out = best_of(lambda: bb.compute_pedersen_commitments(0, cols, g.numpy()))
    ok = all(np.array_equal(refcpu.commit(0, [(cols[c][0][:1 << 12], 0)], g.numpy()[:1 << 12]),
                            bb.compute_pedersen_commitments(0, [(cols[c][0][:1 << 12], 0)], g.numpy()[:1 << 12]))
             for c in (0, 63))
    one = bb.compute_pedersen_commitments(0, [cols[17]], g.numpy())
    print(f"C4 devices={devices}: {dt * 1e3:.2f} ms  {ncols * n / dt:.3e} terms/s (64 x 2^20, one call), "
          f"column 17 equals its single-column call: {np.array_equal(one[0], out[17])}, prefix parity {ok}", flush=True)
if "c5" in which:
    logn = int(os.environ.get("C5_LOGN", "24"))
    n = 1 << logn
    buf = bb.DeviceBuffer(n * 96)
    bb.synthetic_generators_device(2, buf.ptr, n, 0, True)
    p2 = pinned(buf.to_host((n, 96)))
    buf.free()
    t = time.perf_counter()
    h = bb.MultiexpHandle(2, p2.numpy())
    t_new = time.perf_counter() - t
    rng = np.random.default_rng(24)
    sc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x3F
    s = pinned(sc)
    dt, res = best_of(lambda: h.fixed_multiexponentiation(32, 1, n, s.numpy()), iters=3)
    want = common.closed_form_commitment(refcpu, 2, sc)
    ok = common.same(2, refcpu.normalize(2, res), want)
    print(f"C5 devices={devices}: n=2^{logn} {dt * 1e3:.2f} ms  {n / dt:.3e} terms/s (one sxt_fixed_multiexponentiation "
          f"call, host scalars), handle_new {t_new:.2f} s, matches the reference closed form: {ok}", flush=True)
    h.free()
