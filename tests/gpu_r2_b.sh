#!/bin/bash
# round-2 GPU pass B: tables; bench with extras; ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
tail -5 gpurun_out/b_pytest.log
timeout 600 python tests/c5_table_timing.py 21 2 > gpurun_out/b_c5_table.log 2>&1; cat gpurun_out/b_c5_table.log
timeout 600 python tests/c5_table_timing.py 20 0 > gpurun_out/b_c5_table_ed.log 2>&1; cat gpurun_out/b_c5_table_ed.log
( time timeout 900 python bench.py ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -c 6000 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:AccumulateBody -c 2 -o gpurun_out/b_accumulate python tests/prof_c2.py > gpurun_out/b_ncu_full.log 2>&1
tail -3 gpurun_out/b_ncu_full.log
ls -la gpurun_out | tail -12
