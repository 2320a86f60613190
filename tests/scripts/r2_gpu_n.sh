#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/n_pytest.log 2>&1
tail -4 gpurun_out/n_pytest.log
timeout 600 python tests/c5_table_timing.py 20 0 2>&1 | tee gpurun_out/n_table_ed.log
timeout 600 python tests/e2e_c3_ranges.py 1 22 2>&1 | tee gpurun_out/n_c3_ranges.log
timeout 600 python tests/many_columns.py 2>&1 | tail -7 | tee gpurun_out/n_many.log
