#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/pair_timing.py 1 22 0:0 3:32 4:32 4:64 5:32 -1:0 > gpurun_out/d_pair_bls.log 2>&1; cat gpurun_out/d_pair_bls.log
timeout 600 python tests/pair_timing.py 2 21 0:0 2:32 3:32 4:32 4:64 5:32 -1:0 > gpurun_out/d_pair_bn.log 2>&1; cat gpurun_out/d_pair_bn.log
export BLITZAR_B200_PAIR_LEVELS=3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_bn.csv python tests/prof_c2.py 21 1 2 > gpurun_out/d_ncu_bn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:PairPass -c 4 -o gpurun_out/d_pair_bn python tests/prof_c2.py 21 1 2 > gpurun_out/d_ncu_full_bn.log 2>&1
tail -3 gpurun_out/d_ncu_full_bn.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_bls.csv python tests/prof_c2.py 22 1 1 > gpurun_out/d_ncu_bls.log 2>&1
ls -la gpurun_out/d_*
