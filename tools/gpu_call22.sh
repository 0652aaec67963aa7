#!/bin/bash
# conv retune (fp32: 8 rows in flight; bf16: 64-row time tiles): parity tests + bench lines
out=gpurun_out/r02_call22; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 600 python bench.py --no-also > $out/bench_cfg2.json 2> $out/bench_cfg2.err || timeout 600 python bench.py > $out/bench_cfg2.json 2> $out/bench_cfg2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_call22/bench_cfg2.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'])
PY
