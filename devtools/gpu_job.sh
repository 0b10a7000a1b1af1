#!/bin/bash
# Job 54: full GPU test-suite after the launcher change, then the likelihood-kernel captures (traffic.json is keyed on the file hash) and the bench line
mkdir -p gpurun_out
sha256sum sapt_b200/csrc/sapt_loglik_tc2.cu sapt_b200/csrc/sapt_langevin_tc.cu sapt_b200/libsapt.so > gpurun_out/j54_src_sha.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/j54_pytest.txt
NCU="ncu --set full --clock-control none --import-source on"
reduce() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python devtools/prof_kernels.py c3 > gpurun_out/j54_plain_c3.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2j_c3_prof python devtools/prof_kernels.py c3 > gpurun_out/j54_ncu_c3.log 2>&1 && reduce r2j_c3_prof
python devtools/prof_kernels.py c4ll > gpurun_out/j54_plain_c4.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2j_c4_prof python devtools/prof_kernels.py c4ll > gpurun_out/j54_ncu_c4.log 2>&1 && reduce r2j_c4_prof
tail -3 gpurun_out/j54_pytest.txt; tail -1 gpurun_out/j54_ncu_c4.log
