#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
tail -4 gpurun_out/f_pytest.log
for wm in 0 1; do
  echo "== scatter window-major = $wm"
  BLITZAR_B200_SCATTER_WM=$wm timeout 300 python tests/prof_c2.py 20 5 0 2>&1 | tail -2
  BLITZAR_B200_SCATTER_WM=$wm timeout 300 python tests/prof_c2.py 22 4 0 2>&1 | tail -1
  BLITZAR_B200_SCATTER_WM=$wm timeout 300 python tests/prof_c2.py 21 4 2 2>&1 | tail -1
done 2>&1 | tee gpurun_out/f_scatter_wm.log
BLITZAR_B200_SCATTER_WM=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches_wm.csv python tests/prof_c2.py 20 3 0 > /dev/null 2>&1
timeout 600 python tests/ipa_timing.py > gpurun_out/f_ipa.log 2>&1; cat gpurun_out/f_ipa.log
( time timeout 900 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -c 1500 gpurun_out/f_bench.json; tail -5 gpurun_out/f_bench.err
