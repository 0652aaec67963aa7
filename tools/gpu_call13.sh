#!/bin/bash
out=gpurun_out/r02_call13; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stream_fused.py tests/test_gpu_causal.py -x -q 2>&1 | tail -8 | tee $out/tests_stream.log
for b in 1 32 64; do
  timeout 300 python bench.py --workload stream --batch $b --steps 200 --warmup 10 > $out/bench_stream_b$b.json 2> $out/bench_stream_b$b.err
  python -c "
import json
d=json.loads(open('$out/bench_stream_b$b.json').read().strip().splitlines()[-1])
print('b$b', 'ms/push', round(d['ms_per_step'],4), 'audio-s/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" || tail -5 $out/bench_stream_b$b.err
done
timeout 300 python bench.py --workload stream --batch 32 --chunk-ms 10 --steps 200 --warmup 10 > $out/bench_stream_b32_10ms.json 2>/dev/null; python -c "
import json
d=json.loads(open('$out/bench_stream_b32_10ms.json').read().strip().splitlines()[-1]); print('b32 10ms: ms/push', d['ms_per_step'], d['value'])"
timeout 120 python tools/stream_push_timeline.py --frames 20 --batch 32 | cut -c1-1800 | tee $out/timeline_S_b32_f20.json
