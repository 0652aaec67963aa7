#!/bin/bash
out=gpurun_out/r02_call14; mkdir -p $out
for f in "" "--fuse-norm"; do
  tag="cfg3${f:+_fusenorm}"
  timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline $f > $out/bench_$tag.json 2> $out/bench_$tag.err
  python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), round(d['value'],1), d['kernels_ms_per_step'])" || tail -5 $out/bench_$tag.err
done
for f in "" "--fuse-norm"; do
  tag="cfg4${f:+_fusenorm}"
  timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline $f > $out/bench_$tag.json 2> $out/bench_$tag.err
  python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), round(d['value'],1), d['kernels_ms_per_step'])" || tail -5 $out/bench_$tag.err
done
