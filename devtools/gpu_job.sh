#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/j47_pytest.txt
timeout 600 python devtools/x_streaming.py > gpurun_out/j47_xstream.txt 2>&1
timeout 600 python devtools/ab_time.py c3 c4 c5 > gpurun_out/j47_ab.jsonl 2> gpurun_out/j47_ab.err
tail -3 gpurun_out/j47_pytest.txt; cat gpurun_out/j47_xstream.txt; cut -c1-400 gpurun_out/j47_ab.jsonl
