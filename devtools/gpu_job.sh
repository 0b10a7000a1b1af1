#!/bin/bash
mkdir -p gpurun_out
timeout 600 python devtools/ab_time.py c3 c4 acc > gpurun_out/j24_ab.jsonl 2> gpurun_out/j24_ab.err
timeout 900 python -m pytest tests -m gpu -q -x -k "loglik or evaluate or large or full_run or c5 or C5 or fp16 or topolog" 2>&1 | tail -6 > gpurun_out/j24_pytest.txt
cat gpurun_out/j24_ab.jsonl; tail -3 gpurun_out/j24_ab.err; tail -4 gpurun_out/j24_pytest.txt
