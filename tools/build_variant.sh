#!/bin/bash
# Build a variant of ONE source file with extra -D flags and link it against the other objects of the last full
# build: tools/build_variant.sh NAME importance_grp.cu "-DAVR_X=0 ..."  ->  adaptive-volume-rendering_b200/lib/variants/NAME.so
# (A/B experiments on the GPU box: cp the variant over lib/libavr_b200.so inside the gpurun command.)
set -e
cd "$(dirname "$0")/.."
P=adaptive-volume-rendering_b200
mkdir -p $P/lib/variants
name=$1; src=$2; flags=$3
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -cudart static \
  $flags -I include -I $P/csrc -c $P/csrc/$src -o $P/lib/variants/$name.o
objs=$(ls $P/lib/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -Xcompiler -fPIC -o $P/lib/variants/$name.so $objs $P/lib/variants/$name.o
rm $P/lib/variants/$name.o
echo $P/lib/variants/$name.so
