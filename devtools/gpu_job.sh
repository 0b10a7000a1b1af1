#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "swap or sweep or full_run or graph" 2>&1 | tail -8 > gpurun_out/j28_pytest.txt
timeout 300 python devtools/time_rendezvous.py 65536 > gpurun_out/j28_rdv.txt 2>&1
tail -6 gpurun_out/j28_pytest.txt; tail -4 gpurun_out/j28_rdv.txt
