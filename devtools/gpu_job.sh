#!/bin/bash
# The job of the next GPU call (edited between calls; `devtools/gpu.sh <log> <timeout> -- 'bash devtools/gpu_job.sh'`).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/j1_smi.txt 2>&1
python devtools/pipes.py gpurun_out/r2_pipes.json > /dev/null 2> gpurun_out/j1_pipes.err
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_posterior_parity_same_seed_c2 2>&1 | tail -40 > gpurun_out/j1_pytest.txt
for v in "" r1 f2 x16 nco; do
  if [ -z "$v" ]; then lib=sapt_b200/libsapt.so; else lib=sapt_b200/variants/libsapt_$v.so; fi
  SAPT_LIB=$PWD/$lib timeout 600 python devtools/ab_time.py c3 c4 acc >> gpurun_out/j1_ab.jsonl 2>> gpurun_out/j1_ab.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
tail -c 1500 gpurun_out/j1_pytest.txt; cat gpurun_out/j1_ab.jsonl; tail -c 600 gpurun_out/j1_bench.err
