#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/j22_pytest.txt
timeout 600 python devtools/ab_time.py c3 c4 > gpurun_out/j22_ab.jsonl 2> gpurun_out/j22_ab.err
python bench.py --steps 10 --warmup 3 > gpurun_out/j22_bench.json 2> gpurun_out/j22_bench.err
tail -6 gpurun_out/j22_pytest.txt; cat gpurun_out/j22_ab.jsonl; tail -3 gpurun_out/j22_bench.err; cut -c1-400 gpurun_out/j22_bench.json
