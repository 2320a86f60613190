#!/bin/bash
# 2-GPU pass: in-library sharding (BLITZAR_B200_DEVICES), torchrun bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/j_gpus.txt
python - <<'PY' 2>&1 | tee gpurun_out/j_selftest.log
import blitzar_b200 as bb
bb.sxt_init()
print("lane arithmetic selftest mismatches", [bb.selftest_lane_arithmetic(512, s) for s in (1, 2, 3)], flush=True)
PY
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
tail -4 gpurun_out/j_pytest.log
for k in 1 2; do
  BLITZAR_B200_DEVICES=$k C5_LOGN=23 timeout 900 python tests/multi_gpu_abi.py c2 c4 c5
done 2>&1 | tee gpurun_out/j_multi_abi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/j_bench_n2.json 2> gpurun_out/j_bench_n2.err
tail -c 1500 gpurun_out/j_bench_n2.json; tail -5 gpurun_out/j_bench_n2.err
