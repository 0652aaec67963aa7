#!/bin/bash
out=gpurun_out/r02_call18; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_stream_fused.py tests/test_gpu_causal.py -x -q 2>&1 | tail -12 | tee $out/tests.log
for b in 1 8 16 32 64; do
  timeout 300 python bench.py --workload stream --batch $b --steps 200 --warmup 10 > $out/bench_stream_b$b.json 2> $out/bench_stream_b$b.err
  python -c "
import json
d=json.loads(open('$out/bench_stream_b$b.json').read().strip().splitlines()[-1])
print('b$b', 'ms/push', round(d['ms_per_step'],4), 'audio-s/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['latency_ms']['host_observed_median_ms'])" || tail -5 $out/bench_stream_b$b.err
done
timeout 300 python bench.py --workload stream --batch 1 --chunk-ms 2 --steps 200 --warmup 10 > $out/bench_stream_b1_2ms.json 2>/dev/null; python -c "
import json
d=json.loads(open('$out/bench_stream_b1_2ms.json').read().strip().splitlines()[-1]); print('b1 2ms: ms/push', d['ms_per_step'])"
for t in 1 20; do timeout 200 python tools/stack_step_latency.py --tokens $t | tee -a $out/stack_step_latency.jsonl; done
timeout 120 python tools/stream_push_timeline.py --frames 20 | cut -c1-1700 | tee $out/timeline_S_b1_f20.json
