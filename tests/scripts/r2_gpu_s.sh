#!/bin/bash
mkdir -p gpurun_out
for k in 0 1 2; do
  echo "== C2 range skew $k"
  BLITZAR_B200_RANGE_SKEW=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('value %.4e ms %.3f e2e ms %.3f'%(j['value'],j['ms_per_step'],j['e2e']['ms_per_step']))"
done 2>&1 | tee gpurun_out/s_skew_c2.log
for k in 0 -1; do
  echo "== C3 range skew $k"
  PIECES=2,3 BLITZAR_B200_RANGE_SKEW=$k timeout 300 python tests/e2e_c3_ranges.py 1 22 2>&1 | tail -3
done 2>&1 | tee gpurun_out/s_skew_c3.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "upload or default_piece or golden" 2>&1 | tail -2
