#!/bin/bash
# Builds variants of the transform kernels (lab switches in csrc/ntt.cuh) into tools/lab/variants/libapsu_b200_<name>.so;
# a GPU job copies one over apsu_b200/libapsu_b200.so and runs tools/bench_ntt.py.
set -e
cd "$(dirname "$0")/../../apsu_b200/csrc"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
build() { # name, defines
  name=$1; shift
  $NVCC $FLAGS "$@" -c context.cu -o /tmp/context_$name.o 2> ../../tools/lab/variants/$name.ptxas.log
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/lab/variants/libapsu_b200_$name.so /tmp/context_$name.o engine.o capi.o mgpu.o dbbuild.o params.o -cudart static -ldl
  grep -A2 "ntt_kernelILi13ELb[01]ELi32ELi0ELb0" ../../tools/lab/variants/$name.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
}
build u2b3 -DAPSU_NTT_UNROLL=2 -DAPSU_NTT_MINB13=3 &
build u1b2 -DAPSU_NTT_UNROLL=1 -DAPSU_NTT_MINB13=2 &
build u2b2 -DAPSU_NTT_UNROLL=2 -DAPSU_NTT_MINB13=2 &
build u4b2 -DAPSU_NTT_UNROLL=4 -DAPSU_NTT_MINB13=2 &
wait
