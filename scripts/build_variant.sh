#!/bin/bash
# Build a variant of the library that differs in ONE translation unit: scripts/build_variant.sh NAME file.cu -DFLAG=..
# -> casq_b200/lib/variants/libcasq_b200_NAME.so (load it with CASQ_B200_LIB=...).  Tuning aid, not part of the product build.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p casq_b200/lib/variants /tmp/casq_var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c casq_b200/csrc/$src -o /tmp/casq_var_$name/${src%.cu}.o
objs=""
for o in casq_b200/csrc/_obj/*.o; do
  if [ "$(basename $o)" == "${src%.cu}.o" ]; then objs="$objs /tmp/casq_var_$name/${src%.cu}.o"; else objs="$objs $o"; fi
done
nvcc -shared -o casq_b200/lib/variants/libcasq_b200_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart
echo casq_b200/lib/variants/libcasq_b200_$name.so
