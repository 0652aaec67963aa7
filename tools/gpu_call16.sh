#!/bin/bash
# 2-GPU sanity of the driver's command on the final sources (NCCL path, also-array, sequence-parallel parity in the run)
out=gpurun_out/r02_call16; mkdir -p $out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench2 rc=$?"
python -c "
import json
d=json.loads([l for l in open('$out/bench_2gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
for x in d.get('also',[]): print(x['config']['workload'][:30], x['value'], x['ms_per_step'], x['e2e']['ms_per_step'], x.get('parity',{}).get('max_abs_over_rms'), x.get('collectives_per_forward'))" ; tail -3 $out/bench_2gpu.err
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $out/bench_ref_2gpu.json 2> $out/bench_ref_2gpu.err; echo "ref2 rc=$?"; cut -c1-200 $out/bench_ref_2gpu.json
