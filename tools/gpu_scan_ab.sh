#!/bin/bash
# A/B of scan variants on the dev library (tools/devbuild.sh): usage gpu_scan_ab.sh <outdir> <variants>
out=gpurun_out/$1; mkdir -p $out; V=$2
export MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so
timeout 300 python tools/scan_bench.py --variants $V > $out/scan_S_fp32.jsonl 2>&1; cut -c1-200 $out/scan_S_fp32.jsonl
timeout 300 python tools/scan_bench.py --variants $V --mode bf16 > $out/scan_S_bf16.jsonl 2>&1; cut -c1-200 $out/scan_S_bf16.jsonl
timeout 300 python tools/scan_bench.py --variants $V --hparams L --batch 64 --mode bf16 > $out/scan_L_bf16.jsonl 2>&1; cut -c1-200 $out/scan_L_bf16.jsonl
