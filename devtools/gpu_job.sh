#!/bin/bash
# Job 38 (final of round 2): full GPU test-suite, bench line, ncu captures of the shipped kernels (C3 / C4 likelihood, batch-Langevin), launch list
mkdir -p gpurun_out
sha256sum sapt_b200/csrc/sapt_loglik_tc2.cu sapt_b200/csrc/sapt_langevin_tc.cu sapt_b200/libsapt.so > gpurun_out/j38_src_sha.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/j38_pytest.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/j38_bench.json 2> gpurun_out/j38_bench.err
NCU="ncu --set full --clock-control none --import-source on"
reduce() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python devtools/prof_kernels.py c4lt 64 > gpurun_out/j38_plain_lt.log 2>&1 && \
  $NCU -k regex:langevin_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2g_langevin_prof python devtools/prof_kernels.py c4lt 64 > gpurun_out/j38_ncu_lt.log 2>&1 && reduce r2g_langevin_prof
python devtools/prof_kernels.py c3 > gpurun_out/j38_plain_c3.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2g_c3_prof python devtools/prof_kernels.py c3 > gpurun_out/j38_ncu_c3.log 2>&1 && reduce r2g_c3_prof
python devtools/prof_kernels.py c4ll > gpurun_out/j38_plain_c4.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2g_c4_prof python devtools/prof_kernels.py c4ll > gpurun_out/j38_ncu_c4.log 2>&1 && reduce r2g_c4_prof
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/j38_bench_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/j38_ncu_launches.log 2>&1
du -sh gpurun_out; tail -5 gpurun_out/j38_pytest.txt; tail -c 300 gpurun_out/j38_bench.err; cut -c1-300 gpurun_out/j38_bench.json
