#!/bin/bash
# The job of the next GPU call (edited between calls; `devtools/gpu.sh <log> <timeout> -- 'bash devtools/gpu_job.sh'`).
# gpurun_out/ is merged back only while it stays under 64 MiB: ncu reports are reduced to CSV pages on the box.
mkdir -p gpurun_out
sha256sum sapt_b200/csrc/sapt_loglik_tc2.cu sapt_b200/csrc/sapt_langevin_tc.cu sapt_b200/libsapt.so > gpurun_out/j3_src_sha.txt
python devtools/pipes.py gpurun_out/r2_pipes.json > /dev/null 2> gpurun_out/j3_pipes.err
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/j3_pytest.txt
timeout 120 devtools/mma_cost > gpurun_out/j3_mma_cost.txt 2>&1
timeout 600 python devtools/ab_time.py c3 c4 > gpurun_out/j3_ab.jsonl 2> gpurun_out/j3_ab.err
NCU="ncu --set full --clock-control none --import-source on"
reduce() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
python devtools/prof_kernels.py c4lt 64 > gpurun_out/j3_plain_lt.log 2>&1 && \
  $NCU -k regex:langevin_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_langevin_prof python devtools/prof_kernels.py c4lt 64 > gpurun_out/j3_ncu_lt.log 2>&1 && reduce r2_langevin_prof
python devtools/prof_kernels.py c3 > gpurun_out/j3_plain_c3.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2_c3_prof python devtools/prof_kernels.py c3 > gpurun_out/j3_ncu_c3.log 2>&1 && reduce r2_c3_prof
python devtools/prof_kernels.py c4ll > gpurun_out/j3_plain_c4.log 2>&1 && \
  $NCU -k regex:loglik_tc2_kernel -s 1 -c 1 -f -o gpurun_out/r2_c4_prof python devtools/prof_kernels.py c4ll > gpurun_out/j3_ncu_c4.log 2>&1 && reduce r2_c4_prof
python bench.py --steps 10 --warmup 3 > gpurun_out/j3_bench.json 2> gpurun_out/j3_bench.err
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/j3_bench_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu > gpurun_out/j3_ncu_launches.log 2>&1
du -sh gpurun_out; tail -c 1200 gpurun_out/j3_pytest.txt; cat gpurun_out/j3_mma_cost.txt; cat gpurun_out/j3_ab.jsonl; tail -c 400 gpurun_out/j3_bench.err
