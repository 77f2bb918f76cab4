#!/bin/bash
# Instrumented build of the library (per-CTA phase timestamps in score_kernel, scripts/prof_trace.py reads them).
# Never shipped: the product is closyr_b200/libclosyr_b200.so from closyr_b200/csrc/Makefile.
set -e
cd "$(dirname "$0")/../closyr_b200/csrc"
mkdir -p _obj_trace
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2 -DCLOSYR_TRACE"
for f in expr compiler host_api ga; do [ _obj_trace/$f.o -nt $f.cpp ] || $NV -x cu -c $f.cpp -o _obj_trace/$f.o & done
for f in kernels ctx rank; do $NV -c $f.cu -o _obj_trace/$f.o & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libclosyr_b200_trace.so _obj_trace/*.o -cudart static -ldl
echo built ../libclosyr_b200_trace.so
