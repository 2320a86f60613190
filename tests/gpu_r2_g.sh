#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/g_selftest.log
import blitzar_b200 as bb
bb.sxt_init()
for seed in range(1, 6):
    print("lane arithmetic selftest seed", seed, "mismatches", bb.selftest_lane_arithmetic(512, seed), flush=True)
PY
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
tail -4 gpurun_out/g_pytest.log
for lt in 0 1; do
  echo "== lane tail = $lt"
  BLITZAR_B200_LANE_TAIL=$lt timeout 300 python tests/prof_c2.py 20 6 0 2>&1 | tail -2
  BLITZAR_B200_LANE_TAIL=$lt timeout 300 python tests/prof_c2.py 16 6 0 2>&1 | tail -1
done 2>&1 | tee gpurun_out/g_lane_tail.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g_launches.csv python tests/prof_c2.py 20 3 0 > /dev/null 2>&1
timeout 600 python tests/ipa_timing.py > gpurun_out/g_ipa.log 2>&1; cat gpurun_out/g_ipa.log
