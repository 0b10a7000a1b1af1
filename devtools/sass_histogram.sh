#!/bin/bash
# SASS opcode histograms of the tcgen05 kernels in the in-tree libsapt.so (evidence for profiles/: UTCHMMA = tcgen05.mma,
# LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA, FADD2 / FMUL2 = packed fp32 pairs). usage: devtools/sass_histogram.sh > profiles/rN_sass_histogram.txt
cd "$(dirname "$0")/.."
LIB=sapt_b200/libsapt.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  cuobjdump -sass $LIB  (sha256 $(sha256sum $LIB | cut -c1-16))"
for pat in 'loglik_tc2_kernelILi32ELi16ELi1ELi10ELi30E' 'loglik_tc2_kernelILi64ELi8ELi12ELi7E' 'loglik_tc2_kernelILi32ELi32ELi9ELi18E' 'langevin_tc_kernelILi64ELi12E' 'surrogate_tc_kernelILi13E'; do
  fn=$(cuobjdump -sass $LIB 2>/dev/null | grep -o "Function : [^ ]*${pat}[^ ]*" | head -1 | sed 's/Function : //')
  [ -z "$fn" ] && continue
  echo; echo "== $fn"
  cuobjdump -sass -fun "$fn" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed 's/^ *\/\*[0-9a-f]*\*\/ *//' | sed 's/^@!*U*P[0-9T] *//' | awk '{print $1}' | sed 's/;$//' | sed -E 's/^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTCBAR|MUFU\.[A-Z0-9]+|SYNCS)\..*/\1/; s/^([A-Z0-9]+)\..*/\1/' | sort | uniq -c | sort -rn | awk '{printf "%6d %s\n", $1, $2}'
done
