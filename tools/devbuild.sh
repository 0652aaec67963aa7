#!/bin/bash
# Fast-iteration build -> libmtn_b200_dev.so (use with MTN_LIB=... tools/scan_bench.py): recompiles only mtn_scan.cu, restricted to the BASELINE config 2
# instantiation (MTN_SCAN_DEV); all other objects are reused from the last full build.  NOT for shipping: the product library
# libmtn_b200.so is untouched by this script.
set -e
cd "$(dirname "$0")/../avse_challenge_b200/csrc"
nvcc -DMTN_SCAN_DEV -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math \
  -Xcompiler -fPIC -Xptxas -v -c mtn_scan.cu -o mtn_scan_dev.o 2> /tmp/devbuild.log
grep -A2 "scan_kernel_pair" /tmp/devbuild.log | grep -E "registers|spill" || true
nvcc -DMTN_SCAN_DEV -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -c mtn_host.cu -o mtn_host_dev.o
nvcc -shared -o ../libmtn_b200_dev.so mtn_host_dev.o mtn_gemm.o mtn_elem.o mtn_scan_dev.o mtn_seq.o mtn_score.o mtn_dp.o mtn_stream.o mtn_convx.o -lcudart
