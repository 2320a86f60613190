#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r_pytest.log 2>&1
head -3 gpurun_out/r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -c 300 gpurun_out/r_bench.json; tail -3 gpurun_out/r_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:AccumulateBody -c 1 -o gpurun_out/r_accumulate python tests/prof_c2.py 20 1 0 > gpurun_out/r_ncu_full.log 2>&1
timeout 300 python tests/distribution_sweep.py 2>&1 | tee gpurun_out/r_dist.log
timeout 300 python tests/test_ref_gpu_kernels.py 16 18 20 2>&1 | tee gpurun_out/r_refgpu.log
