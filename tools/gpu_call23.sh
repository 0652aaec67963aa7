#!/bin/bash
# GEMM epilogue with warp-uniform control flow (+ 16-warp mask variant): parity tests, micro-benchmark, bench line
out=gpurun_out/r02_call23; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 300 python tools/gemm_bench.py 2>&1 | tee $out/gemm_S_fp32.jsonl | cut -c1-200
timeout 300 python tools/gemm_bench.py --hparams L --batch 64 --mode bf16 2>&1 | tee $out/gemm_L_bf16.jsonl | cut -c1-200
timeout 600 python bench.py --no-also > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_call23/bench_cfg2.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'])
PY
