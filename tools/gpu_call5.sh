#!/bin/bash
out=gpurun_out/r02_call5; mkdir -p $out
timeout 120 python tools/gemm_bench.py --ab --only in_proj > $out/gemm_ab_inproj.jsonl 2>&1; echo "in_proj rc=$?" | tee -a $out/summary.txt; cat $out/gemm_ab_inproj.jsonl | cut -c1-400
timeout 180 python tools/gemm_bench.py --ab > $out/gemm_ab_S_fp32.jsonl 2>&1; echo "S fp32 rc=$?" | tee -a $out/summary.txt; cat $out/gemm_ab_S_fp32.jsonl | cut -c1-330
timeout 180 python tools/gemm_bench.py --ab --hparams L --batch 64 --mode bf16 > $out/gemm_ab_L_bf16.jsonl 2>&1; echo "L bf16 rc=$?" | tee -a $out/summary.txt; cat $out/gemm_ab_L_bf16.jsonl | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm or full_size or small_batch or end_to_end" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt; tail -5 $out/pytest.log
