#!/bin/bash
# A/B builds of the product library: scripts/ab_build.sh <tag> <file.cu> [-DFLAG=..]...  ->  gpurun_ab/libemf_<tag>.so
# (the other objects are taken from emfusion_b200/lib; run `python -m emfusion_b200.build` first)
set -e
tag=$1; src=$2; shift 2
mkdir -p gpurun_ab
obj=gpurun_ab/${src%.cu}_$tag.o
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --default-stream legacy "$@" -c emfusion_b200/csrc/$src -o $obj
others=$(ls emfusion_b200/lib/*.o | grep -v "/${src%.cu}.o" | grep -v safe.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_ab/libemf_$tag.so $obj $others -lcudart
echo gpurun_ab/libemf_$tag.so
