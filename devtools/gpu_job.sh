#!/bin/bash
# The job of the next GPU call (edited between calls; `devtools/gpu.sh <log> <timeout> -- 'bash devtools/gpu_job.sh'`).
mkdir -p gpurun_out
sha256sum sapt_b200/csrc/sapt_loglik_tc2.cu sapt_b200/csrc/sapt_langevin_tc.cu sapt_b200/libsapt.so > gpurun_out/j2_src_sha.txt
python devtools/pipes.py gpurun_out/r2_pipes.json > /dev/null 2> gpurun_out/j2_pipes.err
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/j2_pytest.txt
for v in "" nohs r1; do
  if [ -z "$v" ]; then lib=sapt_b200/libsapt.so; else lib=sapt_b200/variants/libsapt_$v.so; fi
  SAPT_LIB=$PWD/$lib timeout 600 python devtools/ab_time.py c4 acc >> gpurun_out/j2_ab.jsonl 2>> gpurun_out/j2_ab.err
done
SAPT_SURROGATE_TC=0 python devtools/prof_kernels.py c5sp 32768 > gpurun_out/j2_sp.txt 2>&1
SAPT_SURROGATE_TC=1 python devtools/prof_kernels.py c5sp 32768 >> gpurun_out/j2_sp.txt 2>&1
SAPT_SURROGATE_TC=0 python devtools/prof_kernels.py c5sp 4096 >> gpurun_out/j2_sp.txt 2>&1
SAPT_SURROGATE_TC=1 python devtools/prof_kernels.py c5sp 4096 >> gpurun_out/j2_sp.txt 2>&1
timeout 120 devtools/mma_cost > gpurun_out/j2_mma_cost.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
python devtools/prof_kernels.py c4lt 64 > gpurun_out/j2_plain_lt.log 2>&1 && \
  $NCU -k regex:langevin_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_langevin_prof python devtools/prof_kernels.py c4lt 64 > gpurun_out/j2_ncu_lt.log 2>&1
python devtools/prof_kernels.py c3 > gpurun_out/j2_plain_c3.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2_c3_prof python devtools/prof_kernels.py c3 > gpurun_out/j2_ncu_c3.log 2>&1
python devtools/prof_kernels.py c4ll > gpurun_out/j2_plain_c4.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2_c4_prof python devtools/prof_kernels.py c4ll > gpurun_out/j2_ncu_c4.log 2>&1
tail -c 2500 gpurun_out/j2_pytest.txt; cat gpurun_out/j2_ab.jsonl; cat gpurun_out/j2_sp.txt; cat gpurun_out/j2_mma_cost.txt; ls -la gpurun_out/*.ncu-rep | tail -4
