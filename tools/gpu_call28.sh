#!/bin/bash
# final code on 8 GPUs: default bench line (config 2 weak + the also array: config 3 strong, config 5 sequence-parallel)
out=gpurun_out/r02_call28; mkdir -p $out
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_call28/bench_8gpu.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
for a in d.get('also', []): print(a['config']['workload'][:30], a['ms_per_step'], a['value'], a['e2e']['value'], a.get('parity',{}).get('max_abs_over_rms'))
PY
tail -3 $out/bench_8gpu.err
