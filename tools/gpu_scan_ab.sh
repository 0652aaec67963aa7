#!/bin/bash
# One GPU call: v2 scan helper (MTN_SCAN_VARIANT=6) A/B timings, ablations, parity of the scan tests, one ncu capture.
out=gpurun_out/r02_scan2; mkdir -p $out
timeout 300 python tools/scan_bench.py --variants 0,6,S0,S5,S6 > $out/scan_S_fp32.jsonl 2>&1; cat $out/scan_S_fp32.jsonl
MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so timeout 300 python tools/scan_bench.py --variants 6,61,62,63,64,65,66,67,0 > $out/scan_S_fp32_dev.jsonl 2>&1
cat $out/scan_S_fp32_dev.jsonl
MTN_SCAN_VARIANT=6 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scan or end_to_end or golden or chunk" > $out/pytest_variant6.log 2>&1; echo "pytest variant6 rc=$?" | tee -a $out/summary.txt
tail -4 $out/pytest_variant6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_pair2 -s 2 -c 1 -f -o $out/prof_pair2 python tools/scan_bench.py --variants 6 --iters 2 > $out/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $out/summary.txt
tail -3 $out/ncu.log
