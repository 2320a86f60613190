#!/bin/bash
# 8-GPU pass: torchrun bench at N = 8 (and 4), in-library sharding with BLITZAR_B200_DEVICES=8
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/l_gpus.txt
for n in 8 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 ) > gpurun_out/l_bench_n$n.json 2> gpurun_out/l_bench_n$n.err
tail -c 600 gpurun_out/l_bench_n$n.json; tail -3 gpurun_out/l_bench_n$n.err
done
BLITZAR_B200_DEVICES=8 timeout 900 python tests/multi_gpu_abi.py c2 c4 c5 2>&1 | tee gpurun_out/l_multi_abi.log
