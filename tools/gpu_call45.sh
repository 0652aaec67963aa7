#!/bin/bash
# end of round: smoke, whole GPU suite, default bench line, then the ncu launch list of the final code
out=gpurun_out/r02_final3; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/tests.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline > $out/bench_short.json 2>&1; echo "short rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg2_final3.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"; wc -l $out/launches_cfg2_final3.csv
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final3/bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],4), d['roofline']['traffic'], d['cpu_baseline']['value'])
for a in d.get('also', []): print(a['config']['workload'][:26], round(a['ms_per_step'],2), round(a['value']))
PY
