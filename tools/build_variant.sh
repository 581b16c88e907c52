#!/bin/bash
# build a variant of the library with extra -D flags for ONE source (A/B timing on the GPU box via MPB_LIB=...)
# usage: tools/build_variant.sh <tag> <source basename without .cu> <nvcc flags...>
set -e
TAG=$1; SRC=$2; shift 2
D=monopsr_b200/build
mkdir -p $D/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
     --expt-relaxed-constexpr "$@" -c monopsr_b200/csrc/$SRC.cu -o $D/variants/${SRC}_$TAG.o
OBJS=$(ls $D/*.o | grep -v "/$SRC.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $D/variants/lib_$TAG.so $OBJS $D/variants/${SRC}_$TAG.o -lcudart
echo $D/variants/lib_$TAG.so
