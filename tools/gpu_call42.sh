#!/bin/bash
# conv: register ring (bf16 kernel), packed bf16 conversions (fp32 kernel, add_rmsnorm): GPU suite + bench lines
out=gpurun_out/r02_call42; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/tests.log
timeout 900 python bench.py --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_cfg4.json 2>> $out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call42/bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['clocks']['sm_mhz'], round(d['roofline']['frac'],4), d['roofline']['traffic'])
for a in d.get('also', []): print(a['config']['workload'][:26], round(a['ms_per_step'],2), round(a['value']), a.get('parity',{}).get('max_abs_over_rms'))
d=json.loads(open('gpurun_out/r02_call42/bench_cfg4.json').read().strip().splitlines()[-1])
print('cfg4', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['kernels_ms_per_step'].items()})
PY
