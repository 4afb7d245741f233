#!/bin/bash
# Build a profiling variant of the library: tools/build_variant.sh NAME "-DFLAG ..."  ->  build/libob200_NAME.so
set -e
cd "$(dirname "$0")/../optimization_b200/csrc"
name=$1; shift
out=../../build/variant_$name
mkdir -p $out
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in capi tcg_elementwise tcg_stiefel level1 stiefel_tc tcg_stiefel_tc tcg_stiefel_v6 tcg_sphere tcg_sparse lobpcg; do
  [ -f $f.cu ] || continue
  $NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $f.cu -o $out/$f.o &
done
wait
$NVCC $ARCH -shared -o ../../build/libob200_$name.so $out/*.o -lcusolver
echo build/libob200_$name.so
