#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
reduce() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python devtools/prof_kernels.py c5ll > gpurun_out/j41_plain_c5.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2_c5_prof python devtools/prof_kernels.py c5ll > gpurun_out/j41_ncu_c5.log 2>&1 && reduce r2_c5_prof
tail -2 gpurun_out/j41_ncu_c5.log
