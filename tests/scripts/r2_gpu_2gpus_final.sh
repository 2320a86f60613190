#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/p_pytest.log 2>&1
head -3 gpurun_out/p_pytest.log
BLITZAR_B200_DEVICES=2 C5_LOGN=23 timeout 900 python tests/multi_gpu_abi.py c2 c4 c5 2>&1 | tee gpurun_out/p_multi_abi.log
