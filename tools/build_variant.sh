#!/bin/bash
# Build a variant of the library with extra -D flags on ising_fused.cu only: tools/build_variant.sh NAME -DDQ_EXP=1 ...
# Output: variants/lib_NAME.so (git-ignored, travels to the GPU box); use with DIFFQC_B200_LIB=variants/lib_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/../diffquantum_b200/csrc"
mkdir -p ../../variants
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v \
  -c ising_fused.cu -o ../../variants/fused_$name.o 2> ../../variants/fused_$name.ptxas.log
objs=$(ls *.o | grep -v '^ising_fused.o$')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so $objs ../../variants/fused_$name.o
echo built variants/lib_$name.so
