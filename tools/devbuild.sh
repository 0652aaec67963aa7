#!/bin/bash
# Fast-iteration build of libmtn_b200.so: recompiles only mtn_scan.cu, restricted to the BASELINE config 2
# instantiation (MTN_SCAN_DEV); all other objects are reused from the last full build.  NOT for shipping:
# run `python -m avse_challenge_b200.build --force` afterwards.
set -e
cd "$(dirname "$0")/../avse_challenge_b200/csrc"
nvcc -DMTN_SCAN_DEV -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math \
  -Xcompiler -fPIC -Xptxas -v -c mtn_scan.cu -o mtn_scan.o 2> /tmp/devbuild.log
grep -A2 "scan_kernel_pair" /tmp/devbuild.log | grep -E "registers|spill" || true
nvcc -shared -o ../libmtn_b200.so mtn_host.o mtn_gemm.o mtn_elem.o mtn_scan.o mtn_seq.o mtn_score.o mtn_dp.o -lcudart
