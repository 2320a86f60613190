#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1
head -3 gpurun_out/o_pytest.log
timeout 600 python tests/pair_timing.py 1 22 0:0 -1:0 2>&1 | tee gpurun_out/o_pair_bls.log
timeout 600 python tests/pair_timing.py 2 21 0:0 -1:0 2>&1 | tee gpurun_out/o_pair_bn.log
timeout 600 python tests/pair_timing.py 1 16 0:0 -1:0 2:32 2>&1 | tee gpurun_out/o_pair_bls16.log
timeout 600 python tests/pair_timing.py 2 16 0:0 -1:0 2:32 2>&1 | tee gpurun_out/o_pair_bn16.log
timeout 600 python tests/pair_timing.py 1 18 0:0 -1:0 2>&1 | tee gpurun_out/o_pair_bls18.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/o_launches_bls.csv python tests/prof_c2.py 22 1 1 > /dev/null 2>&1
( time timeout 900 python bench.py ) > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
tail -c 300 gpurun_out/o_bench.json; tail -3 gpurun_out/o_bench.err
