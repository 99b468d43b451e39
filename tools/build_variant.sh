#!/bin/bash
# Build a variant of libperfb200.so with extra -D flags into perf_b200/_variants/ (for tools/ab_lib.py):
#   tools/build_variant.sh NAME -DPERF_RS_PAD=16384 ...
name=$1; shift
mkdir -p perf_b200/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
  perf_b200/csrc/*.cu -o perf_b200/_variants/libperfb200_$name.so && echo built perf_b200/_variants/libperfb200_$name.so
