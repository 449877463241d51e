#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=1 ..."  ->  abc_b200/lib/libabc_b200_NAME.so (A/B builds; select with ABC_B200_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
od=/tmp/abc_obj_$name; mkdir -p $od
for f in abc_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c -o $od/$b.o $f &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o abc_b200/lib/libabc_b200_$name.so $od/*.o -lz -ldl
echo built abc_b200/lib/libabc_b200_$name.so
