#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/j36_ab.jsonl
for v in x; do
  timeout 300 python devtools/ab_time.py c4 >> gpurun_out/j36_ab.jsonl 2>> gpurun_out/j36_ab.err
done
cut -c1-200 gpurun_out/j36_ab.jsonl; tail -3 gpurun_out/j36_ab.err
