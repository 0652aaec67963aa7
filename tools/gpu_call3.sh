#!/bin/bash
# GPU call: full GPU suite on the new default scan, dev variants / ablations, default bench (with the also-array), DRAM traffic.
out=gpurun_out/r02_call3; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
grep -E "bf16|passed|failed|Error" $out/pytest_gpu.log | tail -15
MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so timeout 300 python tools/scan_bench.py --variants 0,68,61,62,63,64,65,66,67 > $out/scan_S_fp32_dev.jsonl 2>&1
cut -c1-200 $out/scan_S_fp32_dev.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" | tee -a $out/summary.txt
cut -c1-600 $out/bench_default.json; tail -3 $out/bench_default.err
timeout 600 python tools/scan_traffic.py --out $out/scan_traffic.json > $out/traffic.log 2>&1; tail -4 $out/traffic.log
