"""Sanity + timing of the BASELINE configs on one GPU (device-resident): C2, C3, C4 (per-GPU share:
8 columns), C5 (per-GPU share of the fixed-base MSM)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb
from oracle import port

bb.sxt_init()
rng = np.random.default_rng(0)
which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]

def timed(fn, iters=3):
    best = 1e9
    for _ in range(iters):
        e0, e1 = bb.Event(), bb.Event()
        e0.record(); fn(); e1.record()
        best = min(best, e0.elapsed_ms(e1))
    return best

def tiled_points(curve, n):
    base = port.test_points(curve, 1024, 1)
    reps = n // 1024 + 1
    return np.tile(base[0], (reps, 1))[:n].copy(), np.tile(base[1], (reps, 1))[:n].copy()

if "c2" in which:
    n = 1 << 20
    gens = bb.get_generators(n, 0)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8); s[:, 31] &= 0x0f
    dg, ds, do = bb.DeviceBuffer(host=gens), bb.DeviceBuffer(host=s), bb.DeviceBuffer(64)
    ms = timed(lambda: bb.commit_device(0, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr))
    print(f"C2 ristretto n=2^20 1 col: {ms:.3f} ms  {n/ms*1e3:.3e} terms/s", flush=True)
    for b in (dg, ds, do): b.free()
if "c3" in which:
    n = 1 << 22
    p2, af = tiled_points(1, n)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8); s[:, 31] &= 0x7f
    dg, ds, do = bb.DeviceBuffer(host=af), bb.DeviceBuffer(host=s), bb.DeviceBuffer(64)
    ms = timed(lambda: bb.commit_device(1, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr))
    print(f"C3 bls12-381 n=2^22 1 col: {ms:.3f} ms  {n/ms*1e3:.3e} terms/s", flush=True)
    # parity on a prefix: zero scalars beyond 2^12
    m = 1 << 12
    z = s.copy(); z[m:] = 0
    got = bb.compute_pedersen_commitments(1, [(z, 0)], af)
    print("   prefix parity", np.array_equal(got[:, :48], port.commit(1, [(s[:m], 0)], af[:m])[:, :48]), flush=True)
    # the whole C-ABI call from host memory (upload in pieces) must reproduce the device-resident result
    import torch
    want = do.to_host()[:48].copy() if hasattr(do, "to_host") else None
    for name, pin in (("pageable", False), ("pinned", True)):
        hs, hg = s, af
        if pin:
            hs_t = torch.empty(s.shape, dtype=torch.uint8).pin_memory(); hs_t.numpy()[:] = s; hs = hs_t.numpy()
            hg_t = torch.empty(af.shape, dtype=torch.uint8).pin_memory(); hg_t.numpy()[:] = af; hg = hg_t.numpy()
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); full = bb.compute_pedersen_commitments(1, [(hs, 0)], hg); best = min(best, time.perf_counter() - t)
        ok = want is None or np.array_equal(full[0, :48], want)
        print(f"C3 whole call from {name} host memory: {best*1e3:.1f} ms  {n/best:.3e} terms/s  same as device-resident: {ok}", flush=True)
    for b in (dg, ds, do): b.free()
if "c4" in which:
    n, ncol = 1 << 20, 8
    gens = bb.get_generators(n, 0)
    cols = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(ncol)]
    for c in cols: c[:, 31] &= 0x0f
    dg = bb.DeviceBuffer(host=gens); dss = [bb.DeviceBuffer(host=c) for c in cols]; do = bb.DeviceBuffer(32 * ncol)
    ms = timed(lambda: bb.commit_device(0, [(n, 32, 0)] * ncol, [d.ptr for d in dss], dg.ptr, do.ptr), iters=2)
    print(f"C4 share: ristretto {ncol} cols x 2^20: {ms:.3f} ms  {ncol*n/ms*1e3:.3e} terms/s", flush=True)
    for b in [dg, do] + dss: b.free()
if "c5" in which:
    logn = int(os.environ.get("C5_LOGN", "21"))
    n = 1 << logn
    p2, af = tiled_points(2, n)
    t = time.time(); h = bb.MultiexpHandle(2, p2); print(f"C5 handle_new n=2^{logn}: {time.time()-t:.3f} s", flush=True)
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8); s[:, 31] &= 0x3f
    t = time.time(); res = h.fixed_multiexponentiation(32, 1, n, s); t1 = time.time() - t
    print(f"C5 bn254 fixed n=2^{logn} e2e: {t1*1e3:.1f} ms  {n/t1:.3e} terms/s", flush=True)
    m = 1 << 10
    z = s.copy(); z[m:] = 0
    r2 = h.fixed_multiexponentiation(32, 1, n, z)
    want = port.fixed_msm(2, p2[:m], 1, m, s[:m], element_num_bytes=32)
    print("   prefix parity", np.array_equal(port.normalize(2, r2)[:, :65], port.normalize(2, want)[:, :65]), flush=True)
    h.free()
