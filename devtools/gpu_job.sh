#!/bin/bash
# The job of the next GPU call (edited between calls; `devtools/gpu.sh <log> <timeout> -- 'bash devtools/gpu_job.sh'`).
# Job 4: elected-lane issue of tcgen05.mma / TMA (converged warp + elect.sync) in all three tcgen05 kernels.
mkdir -p gpurun_out
timeout 120 devtools/mma_cost > gpurun_out/j4_mma_cost.txt 2>&1
timeout 600 python devtools/ab_time.py c3 c4 > gpurun_out/j4_ab.jsonl 2> gpurun_out/j4_ab.err
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/j4_pytest.txt
cat gpurun_out/j4_mma_cost.txt; cat gpurun_out/j4_ab.jsonl; tail -5 gpurun_out/j4_ab.err; tail -15 gpurun_out/j4_pytest.txt
