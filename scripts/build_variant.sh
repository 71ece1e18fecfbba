#!/bin/bash
# Perf experiments: build a variant library that differs from libesb.so in the flags of the full-batch step kernel
# (the 16-warp translation unit) only.  usage: scripts/build_variant.sh NAME "-DFLAG=.. ..."  ->  enhancershoebox_b200/variants/libesb_NAME.so
set -e
cd "$(dirname "$0")/../enhancershoebox_b200"
mkdir -p variants build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DESB_TAG=512 -DESB_THREADS=512 -DESB_MIN_CTAS=2 $2 -c csrc/esb_kernels.cu -o build/var_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libesb_$1.so build/esb_api.o build/var_$1.o build/esb_kernels_1024.o build/esb_kernels_2048.o build/esb_aggregate.o
echo built variants/libesb_$1.so
