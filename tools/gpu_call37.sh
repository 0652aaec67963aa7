#!/bin/bash
# residual add in the out_proj epilogue (MTN_EPI_RESADD without planes / row sums, CTA pairs): GPU suite + bench A/B
out=gpurun_out/r02_call37; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/tests.log
for i in 1 2; do
timeout 600 python bench.py --no-also --no-cpu-baseline > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call37/bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, d['clocks']['sm_mhz'])
PY
done
