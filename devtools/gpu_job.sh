#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/j34_ab.jsonl
for v in base c3hc0; do
  SAPT_LIB=$PWD/sapt_b200/variants/libsapt_$v.so timeout 300 python devtools/ab_time.py c3 c5 >> gpurun_out/j34_ab.jsonl 2>> gpurun_out/j34_ab.err
done
cat gpurun_out/j34_ab.jsonl; tail -3 gpurun_out/j34_ab.err
