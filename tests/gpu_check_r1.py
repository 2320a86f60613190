import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import blitzar_b200 as bb
from oracle import refcpu
print('init', bb.sxt_init(num_precomputed_generators=1000))
rng = np.random.default_rng(7)
g = bb.get_generators(5, 3); gr = refcpu.ristretto_generators(5, 3)
print('gens', np.array_equal(refcpu.normalize(0, g), refcpu.normalize(0, gr)))
cols = [[2000,7500,5000,1500],[5000,0,400000,10],[7000,7500,405000,1510]]
columns = [(np.array(c, dtype='<u4').view(np.uint8).reshape(4,4), 0) for c in cols]
o = bb.compute_pedersen_commitments(0, columns); r = refcpu.commit(0, columns)
print('golden', np.array_equal(o, r))
for curve in range(4):
    for n in (300, 5000):
        s = rng.integers(0,256,(n,32),dtype=np.uint8)
        s2 = rng.integers(0,256,(n-7,16),dtype=np.uint8)
        s3 = rng.integers(0,256,(n,8),dtype=np.uint8)
        columns = [(s,0),(s2,1),(s3,1),(s[:1],0),(s[:0],0)]
        gens = refcpu.ristretto_generators(n,0) if curve==0 else refcpu.random_elements(curve,n)[1]
        t=time.time(); o = bb.compute_pedersen_commitments(curve, columns, gens); t1=time.time()-t
        r = refcpu.commit(curve, columns, gens)
        k = 65 if curve > 1 else None
        print(curve, n, [bool(np.array_equal(o[i][:k], r[i][:k])) for i in range(len(columns))], round(t1,3), flush=True)
# timing at 2^20 ristretto (device resident)
for logn in (16, 20):
    n = 1 << logn
    gens = bb.get_generators(n, 0)
    s = rng.integers(0,256,(n,32),dtype=np.uint8); s[:,31] &= 0x0f
    dg = bb.DeviceBuffer(host=gens); ds = bb.DeviceBuffer(host=s); do = bb.DeviceBuffer(64)
    for it in range(3):
        e0, e1 = bb.Event(), bb.Event()
        e0.record()
        bb.commit_device(0, [(n,32,0)], [ds.ptr], dg.ptr, do.ptr)
        e1.record()
        ms = e0.elapsed_ms(e1)
        print('n=2^%d device ms %.3f  terms/s %.3e' % (logn, ms, n/ms*1e3), flush=True)
    t=time.time(); o = bb.compute_pedersen_commitments(0, [(s,0)], gens); print('host-call s', time.time()-t)
    if logn == 16:
        r = refcpu.commit(0, [(s,0)], gens); print('2^16 parity', np.array_equal(o, r), np.array_equal(do.to_host()[:32], r[0]))
    else:
        print('2^20 device vs host-call equal', np.array_equal(do.to_host()[:32], o[0]))
print('launches', bb.launch_count())
