#!/bin/bash
# Build liblotus.so (all sm_100a kernels + the C ABI) in-tree.  Usage: robot-3dlotus_b200/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" \
     -o ../liblotus.so *.cu
