#!/bin/bash
# builds the current csrc/ into exp/<name>/liborlg.so (A/B experiments: the variants are built here, switched on the GPU box)
set -e
name=$1; shift
mkdir -p exp/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas=-v -Xcompiler -fPIC -Xcompiler -pthread -shared -cudart shared "$@" \
  -o exp/$name/liborlg.so optical-rl-gym_b200/csrc/orlg_api.cu > exp/$name/build.log 2>&1
echo built exp/$name
