#!/bin/bash
# final 1-GPU measurement pass: parity, bench, sweep, ncu evidence, sanitizer
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_pytest.log 2>&1
tail -4 gpurun_out/k_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
tail -c 400 gpurun_out/k_bench.json; tail -3 gpurun_out/k_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/k_bench_ref.json 2>&1; tail -c 300 gpurun_out/k_bench_ref.json
timeout 1500 python tests/sweep.py > gpurun_out/k_sweep.md 2>&1; tail -30 gpurun_out/k_sweep.md
timeout 300 python tests/e2e_pageable.py 18 20 22 24 > gpurun_out/k_pageable.log 2>&1; cat gpurun_out/k_pageable.log
timeout 300 python tests/test_ref_gpu_kernels.py 16 18 20 > gpurun_out/k_refgpu.log 2>&1; cat gpurun_out/k_refgpu.log
timeout 600 python tests/ipa_timing.py > gpurun_out/k_ipa.log 2>&1; cat gpurun_out/k_ipa.log
timeout 600 python tests/distribution_sweep.py > gpurun_out/k_dist.log 2>&1; cat gpurun_out/k_dist.log
timeout 900 python tests/many_columns.py > gpurun_out/k_many.log 2>&1; cat gpurun_out/k_many.log
# ncu: launch list of the bench command; full capture of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/k_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:AccumulateBody -c 1 -o gpurun_out/k_accumulate python tests/prof_c2.py 20 1 0 > gpurun_out/k_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches_bls.csv python tests/prof_c2.py 22 1 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:PairPass2 -c 1 -o gpurun_out/k_pair_bls python tests/prof_c2.py 22 1 1 > gpurun_out/k_ncu_full_bls.log 2>&1
bash tests/scripts/sanitizer.sh
ls -la gpurun_out/k_* gpurun_out/sanitizer_*
