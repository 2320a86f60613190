#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m | head -8
python - <<'PY' 2>&1 | tail -40
import os, sys, time
os.environ["BLITZAR_B200_DEVICES"]="2"
os.environ["BLITZAR_B200_TRACE"]="1"
import numpy as np, torch
sys.path.insert(0, ".")
import blitzar_b200.api as bb
from tests import common
bb.sxt_init()
n=1<<20
def pinned(a):
    t=torch.empty(a.shape,dtype=torch.uint8).pin_memory(); t.numpy()[:]=a; return t
g=pinned(bb.get_generators(n,0)); s=pinned(common.mt19937_bytes(0,n))
for it in range(8):
    t=time.perf_counter(); out=bb.compute_pedersen_commitments(0,[(s.numpy(),0)],g.numpy()); dt=time.perf_counter()-t
    print("iter",it,"ms %.2f"%(dt*1e3),flush=True)
PY
