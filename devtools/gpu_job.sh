#!/bin/bash
# Job 52: ncu captures of the likelihood kernel after the last edit of its source file (traffic.json is keyed on the file hash), X-streaming numbers
mkdir -p gpurun_out
sha256sum sapt_b200/csrc/sapt_loglik_tc2.cu sapt_b200/csrc/sapt_langevin_tc.cu sapt_b200/libsapt.so > gpurun_out/j52_src_sha.txt
NCU="ncu --set full --clock-control none --import-source on"
reduce() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python devtools/prof_kernels.py c3 > gpurun_out/j52_plain_c3.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2i_c3_prof python devtools/prof_kernels.py c3 > gpurun_out/j52_ncu_c3.log 2>&1 && reduce r2i_c3_prof
python devtools/prof_kernels.py c4ll > gpurun_out/j52_plain_c4.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2i_c4_prof python devtools/prof_kernels.py c4ll > gpurun_out/j52_ncu_c4.log 2>&1 && reduce r2i_c4_prof
timeout 600 python devtools/x_streaming.py > gpurun_out/j52_xstream.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/j52_bench.json 2> gpurun_out/j52_bench.err
tail -2 gpurun_out/j52_ncu_c4.log; head -4 gpurun_out/j52_xstream.txt
