"""Inner-product argument timing on the GPU box (prove / verify through the C ABI) beside the
reference cpu backend on a smaller n."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blitzar_b200 as bb
from oracle import refcpu
L = 2**252 + 27742317777372353535851937790883648493
bb.sxt_init(num_precomputed_generators=(1 << 18) + 1)
rng = np.random.default_rng(0)
def scal(n):
    x = rng.integers(0, 256, (n, 32), dtype=np.uint8); x[:, 31] &= 0x0f
    return x
for logn in (10, 14, 16, 18):
    n = 1 << logn
    a, b = scal(n), scal(n)
    t0 = refcpu.transcript_new(b"timing")
    t = t0.copy(); bb.prove_inner_product(t, a, b, 0)
    t = t0.copy(); s = time.perf_counter(); lv, rv, ap = bb.prove_inner_product(t, a, b, 0); dt = time.perf_counter() - s
    line = f"n=2^{logn}: prove {dt*1e3:.1f} ms"
    if logn <= 10:
        tr = t0.copy(); s = time.perf_counter(); want = refcpu.prove_inner_product(tr, a, b, 0); dr = time.perf_counter() - s
        line += f" (reference cpu {dr*1e3:.1f} ms, same proof: {all(np.array_equal(x, y) for x, y in zip(want, (lv, rv, ap)))})"
    # verify: product = <a, b> mod l, a_commit = <a, G> (one MSM through the commitments API)
    ai = [int.from_bytes(bytes(r), "little") for r in a]
    bi = [int.from_bytes(bytes(r), "little") for r in b]
    prod = np.frombuffer((sum(x * y for x, y in zip(ai, bi)) % L).to_bytes(32, "little"), dtype=np.uint8)
    gens = bb.get_generators(n, 0)
    dg, ds = bb.DeviceBuffer(host=gens), bb.DeviceBuffer(host=a)
    part, out = bb.DeviceBuffer(128), bb.DeviceBuffer(160)
    bb.commit_device(0, [(n, 32, 0)], [ds.ptr], dg.ptr, None, part.ptr)
    bb.combine_partials_projective_device(0, out.ptr, part.ptr, 1, 1)
    acommit = out.to_host()[:160].copy()
    t = t0.copy(); bb.verify_inner_product(t, b, prod, acommit, lv, rv, ap, 0)
    t = t0.copy(); s = time.perf_counter(); ok = bb.verify_inner_product(t, b, prod, acommit, lv, rv, ap, 0); dv = time.perf_counter() - s
    line += f", verify {dv*1e3:.1f} ms (accepted: {ok == 1})"
    print(line, flush=True)
