#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1
head -3 gpurun_out/t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tests/prof_c2.py 20 6 0 2>&1 | tail -2
timeout 300 python tests/prof_c2.py 22 3 1 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/t_bench.json 2>/dev/null
python -c "import json; j=json.loads(open('gpurun_out/t_bench.json').readline()); print('value %.4e ms %.3f e2e ms %.3f share %.3f'%(j['value'],j['ms_per_step'],j['e2e']['ms_per_step'],j['roofline']['kernel_share_of_step']))"
