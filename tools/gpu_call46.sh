#!/bin/bash
out=gpurun_out/r02_final5; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/tests.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final5/bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],4), d['roofline']['traffic'], d['cpu_baseline']['value'])
for a in d.get('also', []): print(a['config']['workload'][:26], round(a['ms_per_step'],2), round(a['value']))
PY
