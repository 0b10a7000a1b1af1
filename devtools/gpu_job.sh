#!/bin/bash
mkdir -p gpurun_out
python devtools/c4_steps.py 4 > gpurun_out/j57_c4_plain.log 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c4_launches.csv python devtools/c4_steps.py 4 > gpurun_out/j57_c4_ncu.log 2>&1
tail -2 gpurun_out/j57_c4_plain.log
