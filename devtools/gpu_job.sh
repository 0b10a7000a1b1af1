#!/bin/bash
mkdir -p gpurun_out
timeout 300 python devtools/lt_check.py > gpurun_out/j15_lt_check.txt 2>&1
timeout 300 python devtools/lt_check.py 40-64-7 34-30-16 60-20-3 16-16-2 > gpurun_out/j15_lt_check2.txt 2>&1
SAPT_LT_TRACE=1 timeout 300 python devtools/prof_kernels.py c4lt 64 > gpurun_out/j15_trace.txt 2>&1
timeout 600 python devtools/ab_time.py c4 > gpurun_out/j15_ab.jsonl 2> gpurun_out/j15_ab.err
grep -hv "^   " gpurun_out/j15_lt_check.txt gpurun_out/j15_lt_check2.txt; cat gpurun_out/j15_ab.jsonl
