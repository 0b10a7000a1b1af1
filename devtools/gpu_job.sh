#!/bin/bash
# Job 6: batch-Langevin kernel as two independent tile streams (3 X stages, bias folded, polled MMA state machine)
mkdir -p gpurun_out
timeout 300 python devtools/lt_check.py > gpurun_out/j6_lt_check.txt 2>&1
timeout 300 python devtools/lt_check.py 40-64-7 34-30-16 60-20-3 16-16-2 > gpurun_out/j6_lt_check2.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "langevin or lg or c4 or full_run or graph" 2>&1 | tail -15 > gpurun_out/j6_pytest.txt
timeout 600 python devtools/ab_time.py c4 > gpurun_out/j6_ab.jsonl 2> gpurun_out/j6_ab.err
cat gpurun_out/j6_lt_check.txt gpurun_out/j6_lt_check2.txt | grep -v "^   "; tail -8 gpurun_out/j6_pytest.txt; cat gpurun_out/j6_ab.jsonl; tail -3 gpurun_out/j6_ab.err
