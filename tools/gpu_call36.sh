#!/bin/bash
# 256-byte L2 promotion for the A operand of the narrow-N fp32-mode GEMMs (x_proj): full GPU suite, GEMM micro-benchmark, bench line
out=gpurun_out/r02_call36; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/tests.log
timeout 300 python tools/gemm_bench.py --only x_proj 2>&1 | cut -c1-200
timeout 600 python bench.py --no-also --no-cpu-baseline > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call36/bench_cfg2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'])
PY
