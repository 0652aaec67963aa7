#!/bin/bash
# final validation on a forced rebuild: smoke(), full GPU suite, default bench line, reference arm
out=gpurun_out/r02_call40; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/tests.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call40/bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'])
r=json.loads(open('gpurun_out/r02_call40/bench_reference.json').read().strip().splitlines()[-1])
print(r['impl'], r['value'], r['cpu_baseline']['sample'][:80])
PY
