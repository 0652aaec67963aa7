#!/bin/bash
# final-sources evidence: whole GPU suite, smoke(), default bench, launch list
out=gpurun_out/r02_call15; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$out/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print(d['kernels_ms_per_step'])
for x in d.get('also',[]): print(x['config']['workload'][:30], x['value'], x['ms_per_step'], x.get('parity'))"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-400 $out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 250 --csv --log-file $out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
