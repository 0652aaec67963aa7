#!/bin/bash
# decoder frames with 4 rows per warp: parity tests + bench line
out=gpurun_out/r02_call33; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/tests.log
timeout 600 python bench.py --no-also --no-cpu-baseline > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_call33/bench_cfg2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'])
PY
