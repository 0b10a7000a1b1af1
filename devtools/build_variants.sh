#!/bin/bash
# A/B builds of libsapt.so: recompiles sapt_loglik_tc2.cu with different -D switches and relinks with the other objects.
# usage: devtools/build_variants.sh   -> sapt_b200/variants/libsapt_<name>.so   (select with SAPT_LIB=<path>)
set -e
cd "$(dirname "$0")/../sapt_b200/csrc"
make -j8 > /dev/null
mkdir -p ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I."
OTHERS=$(ls *.o | grep -v sapt_loglik_tc2.o)
build() { name=$1; shift
  nvcc $FLAGS "$@" -c sapt_loglik_tc2.cu -o ../variants/tc2_$name.o
  nvcc $ARCH -shared -o ../variants/libsapt_$name.so $OTHERS ../variants/tc2_$name.o -lcuda
  echo built $name; }
build r1   -DSAPT_F32X2=0 -DSAPT_X_FP16=0 -DSAPT_CONST_O=0 -DSAPT_HS_SMEM=0 &
build nohs -DSAPT_HS_SMEM=0 &
wait
