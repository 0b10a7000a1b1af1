#!/bin/bash
# A/B builds of libsapt.so: recompiles sapt_loglik_tc2.cu with different -D switches and relinks with the other objects.
# usage: devtools/build_variants.sh name1 "flags1" name2 "flags2" ...   -> sapt_b200/variants/libsapt_<name>.so   (select with SAPT_LIB=<path>)
# (development builds instantiate only the BASELINE shapes: -DSAPT_TC2_FEW)
set -e
cd "$(dirname "$0")/../sapt_b200/csrc"
make -j16 > /dev/null
mkdir -p ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. -DSAPT_TC2_FEW"
OTHERS=$(ls *.o | grep -v sapt_loglik_tc2.o)
build() { name=$1; shift
  nvcc $FLAGS $@ -c sapt_loglik_tc2.cu -o ../variants/tc2_$name.o
  nvcc $ARCH -shared -o ../variants/libsapt_$name.so $OTHERS ../variants/tc2_$name.o -lcuda
  echo built $name; }
while [ $# -gt 1 ]; do build "$1" $2 & shift; shift; done
wait
