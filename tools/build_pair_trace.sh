#!/bin/sh
# development build of the library with the CTA-pair contraction's per-tile trace (tools/trace_pair.py); the product build is untouched
set -e
cd "$(dirname "$0")/../adaptive_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr \
     -DAA_PAIR_TRACE -c gemm_tc.cu -o build/gemm_tc_trace.o
objs=$(ls build/*.o | grep -v "gemm_tc.o" | grep -v "gemm_tc_trace.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/libadaptive_trace.so $objs build/gemm_tc_trace.o
