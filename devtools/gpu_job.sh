#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/j35_ab.jsonl
for v in base sd1; do
  SAPT_LIB=$PWD/sapt_b200/variants/libsapt_$v.so timeout 300 python devtools/ab_time.py c4 >> gpurun_out/j35_ab.jsonl 2>> gpurun_out/j35_ab.err
done
cut -c1-200 gpurun_out/j35_ab.jsonl; tail -3 gpurun_out/j35_ab.err
