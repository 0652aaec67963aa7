#!/bin/bash
# residual add in the out_proj epilogue for every driver (engine, MambaStack / DPMamba, sequence-parallel backend)
out=gpurun_out/r02_call38; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/tests.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
timeout 600 python bench.py --model dpmamba --hparams S --workload custom --no-cpu-baseline --steps 10 --warmup 3 > $out/bench_dpmamba_S.json 2>> $out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call38/bench_default.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))
for a in d.get('also', []): print(a['config']['workload'][:26], round(a['ms_per_step'],2), round(a['value']), a.get('parity',{}).get('max_abs_over_rms'))
d=json.loads(open('gpurun_out/r02_call38/bench_dpmamba_S.json').read().strip().splitlines()[-1])
print('dpmamba S', round(d['ms_per_step'],2), round(d['value']))
PY
