#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/j42_ab.jsonl
for v in w0 w20 w50 w100; do
  SAPT_LIB=$PWD/sapt_b200/variants/libsapt_$v.so timeout 300 python devtools/ab_time.py c3 c4 c5 >> gpurun_out/j42_ab.jsonl 2>> gpurun_out/j42_ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/j42_ab.jsonl'):
    j=json.loads(l); print(j['lib'], 'c3', j['c3_loglik_ms'][0], 'c4', j['c4_loglik_ms'][0], 'c5', j['c5_loglik_ms'][0])
PY
tail -3 gpurun_out/j42_ab.err
