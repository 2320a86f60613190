"""Throughput sweep n = 2^16 .. 2^24: device-resident kernel time and end-to-end
C-ABI time per curve / API on one GPU. Writes gpurun_out/sweep.json and prints a markdown table."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb
from oracle import port

bb.sxt_init()
rng = np.random.default_rng(0)
rows = []

def timed(fn, iters=3):
    best = 1e9
    for _ in range(iters):
        e0, e1 = bb.Event(), bb.Event()
        e0.record(); fn(); e1.record()
        best = min(best, e0.elapsed_ms(e1))
    return best

def wall(fn, iters=3):
    best = 1e9
    for _ in range(iters):
        t = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t)
    return best * 1e3

def tiled(curve, n):
    p2, af = port.test_points(curve, 1024, 1)
    reps = n // 1024 + 1
    return np.tile(p2, (reps, 1))[:n].copy(), np.tile(af, (reps, 1))[:n].copy()

max_log = int(os.environ.get("SWEEP_MAX", "24"))
for logn in range(16, max_log + 1, 2):
    n = 1 << logn
    s = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    # ristretto var-base
    gens = bb.get_generators(n, 0)
    s0 = s.copy(); s0[:, 31] &= 0x0f
    dg, ds, do = bb.DeviceBuffer(host=gens), bb.DeviceBuffer(host=s0), bb.DeviceBuffer(64)
    ms = timed(lambda: bb.commit_device(0, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr))
    e2e = wall(lambda: bb.compute_pedersen_commitments(0, [(s0, 0)], gens))
    rows.append(dict(path="ristretto255 var-base", logn=logn, kernel_ms=ms, e2e_ms=e2e, bytes_per_term=192))
    for b in (dg, ds, do): b.free()
    del gens
    if logn <= 22:
        for curve, name, mask, bpt in ((1, "bls12-381 var-base", 0x7f, 136), (2, "bn254 var-base", 0x3f, 104)):
            p2, af = tiled(curve, n)
            sc = s.copy(); sc[:, 31] &= mask
            dg, ds, do = bb.DeviceBuffer(host=af), bb.DeviceBuffer(host=sc), bb.DeviceBuffer(128)
            ms = timed(lambda: bb.commit_device(curve, [(n, 32, 0)], [ds.ptr], dg.ptr, do.ptr))
            e2e = wall(lambda: bb.compute_pedersen_commitments(curve, [(sc, 0)], af), iters=2)
            rows.append(dict(path=name, logn=logn, kernel_ms=ms, e2e_ms=e2e, bytes_per_term=bpt))
            for b in (dg, ds, do): b.free()
    # bn254 fixed-base (handle resident; e2e = H2D scalars + MSM + D2H)
    p2, af = tiled(2, n)
    sc = s.copy(); sc[:, 31] &= 0x3f
    t = time.perf_counter(); h = bb.MultiexpHandle(2, p2); t_handle = time.perf_counter() - t
    e2e = wall(lambda: h.fixed_multiexponentiation(32, 1, n, sc), iters=2)
    rows.append(dict(path="bn254 fixed-base (handle)", logn=logn, kernel_ms=None, e2e_ms=e2e, bytes_per_term=96, handle_s=t_handle))
    h.free()
    print("done 2^%d" % logn, flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)
print("| path | n | kernel ms | terms/s (kernel) | GB/s algorithmic | e2e ms | terms/s (e2e) |")
print("|---|---|---|---|---|---|---|")
for r in rows:
    n = 1 << r["logn"]
    k = r["kernel_ms"]
    print(f"| {r['path']} | 2^{r['logn']} | {k:.3f} | {n/k*1e3:.3e} | {r['bytes_per_term']*n/k/1e6:.1f} | {r['e2e_ms']:.2f} | {n/r['e2e_ms']*1e3:.3e} |" if k else
          f"| {r['path']} | 2^{r['logn']} | - | - | - | {r['e2e_ms']:.2f} | {n/r['e2e_ms']*1e3:.3e} |")
