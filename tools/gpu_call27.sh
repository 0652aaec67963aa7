#!/bin/bash
# forward_host (pipelined host API): tests + default bench line with the also array
out=gpurun_out/r02_call27; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_call27/bench_default.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'], d['roofline']['frac'])
for a in d.get('also', []): print(a['config']['workload'][:30], a['ms_per_step'], a['value'], a['e2e']['value'])
PY
tail -3 $out/bench_default.err
