#!/bin/bash
out=gpurun_out/r02_call6; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
grep -E "bf16|fp32 B=|passed|failed|FAILED" $out/pytest_gpu.log | tail -20
MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so timeout 300 python tools/scan_bench.py --variants 0,69 --mode bf16 > $out/scan_S_bf16_softplus.jsonl 2>&1; cut -c1-220 $out/scan_S_bf16_softplus.jsonl
MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so timeout 300 python tools/scan_bench.py --variants 0,69 --hparams L --batch 64 --mode bf16 > $out/scan_L_bf16_softplus.jsonl 2>&1; cut -c1-220 $out/scan_L_bf16_softplus.jsonl
MTN_LIB=avse_challenge_b200/libmtn_b200_dev.so timeout 300 python tools/scan_bench.py --variants 0,69 > $out/scan_S_fp32_softplus.jsonl 2>&1; cut -c1-220 $out/scan_S_fp32_softplus.jsonl
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" | tee -a $out/summary.txt
cut -c1-300 $out/bench_default.json; tail -3 $out/bench_default.err
timeout 600 python tools/scan_traffic.py --out $out/scan_traffic.json > $out/traffic.log 2>&1; tail -4 $out/traffic.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 250 --csv --log-file $out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu rc=$?" | tee -a $out/summary.txt
