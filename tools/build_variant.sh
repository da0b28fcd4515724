#!/bin/bash
# build a kernel-variant copy of the product library for A/B experiments: tools/build_variant.sh NAME -DIFL_SW_U=4 ...
set -e
name=$1; shift
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../incremental-fluids-java_b200/csrc
obj=$(mktemp -d)
for f in ifl_api ifl_grid_kernels ifl_solve_kernels ifl_solid_kernels ifl_flip_kernels ifl_comm; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c $src/$f.cu -o $obj/$f.o &
done
wait
mkdir -p $here/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $here/_variants/lib_$name.so $obj/*.o -ldl
rm -rf $obj
