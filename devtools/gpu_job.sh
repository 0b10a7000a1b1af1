#!/bin/bash
mkdir -p gpurun_out
python devtools/c5_steps.py 8 > gpurun_out/j31_c5_plain.log 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c5_launches.csv python devtools/c5_steps.py 8 > gpurun_out/j31_c5_ncu.log 2>&1
tail -2 gpurun_out/j31_c5_plain.log
