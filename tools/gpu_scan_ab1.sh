#!/bin/bash
# A/B of scan variants on the dev library, config-2 shape only: usage gpu_scan_ab1.sh <outdir> <variants>
out=gpurun_out/$1; mkdir -p $out; V=$2
export MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so
timeout 300 python tools/scan_bench.py --variants $V > $out/scan_S_fp32.jsonl 2>&1; cut -c1-130 $out/scan_S_fp32.jsonl
