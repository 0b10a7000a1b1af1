#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/j29_ab.jsonl
for v in h2s0 h2s8 h2s5; do
  SAPT_LIB=$PWD/sapt_b200/variants/libsapt_$v.so timeout 300 python devtools/ab_time.py c3 acc >> gpurun_out/j29_ab.jsonl 2>> gpurun_out/j29_ab.err
done
SAPT_LIB=$PWD/sapt_b200/variants/libsapt_trh2s8.so timeout 300 python devtools/prof_kernels.py c3 > gpurun_out/j29_trace_c3.txt 2>&1
cut -c1-330 gpurun_out/j29_ab.jsonl; tail -3 gpurun_out/j29_ab.err
