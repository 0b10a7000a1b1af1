#!/bin/bash
mkdir -p gpurun_out
timeout 300 python devtools/lt_check.py > gpurun_out/j30_lt_check.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "langevin or lg or c4" 2>&1 | tail -6 > gpurun_out/j30_pytest.txt
timeout 600 python devtools/ab_time.py c4 > gpurun_out/j30_ab.jsonl 2> gpurun_out/j30_ab.err
grep -hv "^   " gpurun_out/j30_lt_check.txt; tail -4 gpurun_out/j30_pytest.txt; cat gpurun_out/j30_ab.jsonl
