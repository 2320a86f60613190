#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
tail -5 gpurun_out/c_pytest.log
timeout 600 python tests/pair_timing.py 1 22 > gpurun_out/c_pair_bls.log 2>&1; cat gpurun_out/c_pair_bls.log
timeout 600 python tests/pair_timing.py 2 21 > gpurun_out/c_pair_bn.log 2>&1; cat gpurun_out/c_pair_bn.log
timeout 600 python tests/c5_table_timing.py 21 2 > gpurun_out/c_c5_table.log 2>&1; cat gpurun_out/c_c5_table.log
