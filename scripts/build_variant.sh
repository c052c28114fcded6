#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: an alternative build of the library under variants/<name>.so
# (A/B timing with PERCNN_B200_LIB=variants/<name>.so); every TU is recompiled with the extra flags.
set -e
name=$1; shift
mkdir -p variants/obj_$name
for f in percnn_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o variants/obj_$name/$b.o $f &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o variants/$name.so variants/obj_$name/*.o
rm -rf variants/obj_$name
echo built variants/$name.so
