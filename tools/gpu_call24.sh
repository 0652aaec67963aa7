#!/bin/bash
# GEMM epilogue rewrite + 16-warp variants (fp32 mask, bf16 in_proj / mask): full GPU suite, config 3 / 4 bench lines
out=gpurun_out/r02_call24; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_cfg3.json 2> $out/bench_cfg3.err
timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_cfg4.json 2> $out/bench_cfg4.err
python - <<'PY'
import json
for c in ('cfg3','cfg4'):
    d=json.load(open(f'gpurun_out/r02_call24/bench_{c}.json'))
    print(c, d['ms_per_step'], d['value'], d['kernels_ms_per_step'], d['clocks'], d['roofline']['frac'])
PY
