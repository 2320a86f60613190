#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/tune.py 20 0 > gpurun_out/m_tune_c2.log 2>&1
sort -t: -k2 -n gpurun_out/m_tune_c2.log | awk '{print}' | sort -k6 -n | head -12
echo ...; grep "c=0 k1=0 kn=8 g1=16 gn=4" gpurun_out/m_tune_c2.log
timeout 300 python tests/tune.py 16 0 0,0,8,16,4 0,0,8,8,4 0,0,4,8,4 0,0,8,8,2 0,0,8,4,4 > gpurun_out/m_tune_2_16.log 2>&1; cat gpurun_out/m_tune_2_16.log
