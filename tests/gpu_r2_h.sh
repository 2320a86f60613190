#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/h_selftest.log
import blitzar_b200 as bb
bb.sxt_init()
for seed in range(1, 4):
    print("lane arithmetic selftest seed", seed, "mismatches", bb.selftest_lane_arithmetic(512, seed), flush=True)
PY
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/h_pytest.log 2>&1
tail -4 gpurun_out/h_pytest.log
for t in 2 4 8 16; do BLITZAR_B200_STAGER_THREADS=$t timeout 300 python tests/e2e_pageable.py 20 22; done 2>&1 | tee gpurun_out/h_pageable.log
( time timeout 900 python bench.py ) > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -c 600 gpurun_out/h_bench.json; tail -3 gpurun_out/h_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/h_ncu_bench.log 2>&1
