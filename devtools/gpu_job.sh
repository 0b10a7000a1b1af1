#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/j56_pytest.txt
tail -2 gpurun_out/j56_pytest.txt
