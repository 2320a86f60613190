#!/bin/bash
# round-2 GPU pass A: parity (incl. BASELINE sizes), bench with extras, ncu of the accumulate kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:AccumulateBody -c 2 -o gpurun_out/a_accumulate python tests/prof_c2.py > gpurun_out/a_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
