#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1
tail -5 gpurun_out/e_pytest.log
timeout 600 python tests/pair_timing.py 1 22 0:0 2:32 3:32 4:32 -1:0 > gpurun_out/e_pair_bls.log 2>&1; cat gpurun_out/e_pair_bls.log
timeout 600 python tests/pair_timing.py 2 21 0:0 1:32 2:32 3:32 -1:0 > gpurun_out/e_pair_bn.log 2>&1; cat gpurun_out/e_pair_bn.log
timeout 600 python tests/c5_table_timing.py 21 2 > gpurun_out/e_c5_table.log 2>&1; cat gpurun_out/e_c5_table.log
timeout 600 python tests/ipa_timing.py > gpurun_out/e_ipa.log 2>&1; cat gpurun_out/e_ipa.log
timeout 600 python tests/distribution_sweep.py > gpurun_out/e_dist.log 2>&1; cat gpurun_out/e_dist.log
timeout 900 python tests/many_columns.py > gpurun_out/e_many.log 2>&1; cat gpurun_out/e_many.log
