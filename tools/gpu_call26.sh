#!/bin/bash
# encoder (4 tokens per warp iteration) + decoder frames (pair rows, transposing reduce): parity tests + bench line
out=gpurun_out/r02_call26; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 600 python bench.py --no-also > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_call26/bench_cfg2.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'], d['roofline']['frac'])
PY
