#!/bin/bash
mkdir -p gpurun_out
for u in 0 1; do
  echo "== uniform add = $u"
  BLITZAR_B200_UNIFORM_ADD=$u timeout 300 python tests/prof_c2.py 20 6 0 2>&1 | tail -2
  BLITZAR_B200_UNIFORM_ADD=$u timeout 300 python tests/prof_c2.py 22 4 0 2>&1 | tail -1
  BLITZAR_B200_UNIFORM_ADD=$u timeout 300 python tests/prof_c2.py 16 6 0 2>&1 | tail -1
  BLITZAR_B200_UNIFORM_ADD=$u BLITZAR_B200_PAIR_LEVELS=0 timeout 300 python tests/prof_c2.py 21 4 2 2>&1 | tail -1
  BLITZAR_B200_UNIFORM_ADD=$u timeout 300 python tests/prof_c2.py 22 3 1 2>&1 | tail -1
done 2>&1 | tee gpurun_out/q_uniform.log
BLITZAR_B200_UNIFORM_ADD=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/q_launches_uniform.csv python tests/prof_c2.py 20 2 0 > /dev/null 2>&1
BLITZAR_B200_UNIFORM_ADD=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
