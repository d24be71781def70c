#!/bin/bash
# Build a tuning variant of the CUDA library: tools/build_variant.sh <out.so> <extra nvcc flags...>
set -e
out=$1; shift
cd "$(dirname "$0")/.."
unset CC CXX
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC,-O2 -shared -cudart static \
  -I include -I landhydrology.jl_b200/csrc "$@" -o "$out" -t 4 landhydrology.jl_b200/csrc/*.cu -ldl
