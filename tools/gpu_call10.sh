#!/bin/bash
# 1-GPU call: whole GPU test suite on the current build + streaming bench lines (fused push vs batch plan).
out=gpurun_out/r02_call10; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/gpu_tests.log
for b in 1 32; do
  for f in "" "--no-fused-push"; do
    tag="b${b}${f:+_batchplan}"
    timeout 300 python bench.py --workload stream --batch $b --steps 200 --warmup 10 $f > $out/bench_stream_$tag.json 2> $out/bench_stream_$tag.err
    echo "stream $tag rc=$?"; python -c "
import json,sys
d=json.loads(open('$out/bench_stream_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'ms/push', round(d['ms_per_step'],4), 'audio-s/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['latency_ms'])
" || tail -5 $out/bench_stream_$tag.err
  done
done
timeout 300 python bench.py --workload stream --batch 1 --chunk-ms 2 --steps 200 --warmup 10 > $out/bench_stream_b1_2ms.json 2>/dev/null; python -c "
import json
d=json.loads(open('$out/bench_stream_b1_2ms.json').read().strip().splitlines()[-1]); print('2ms chunks: ms/push', d['ms_per_step'], d['latency_ms'])"
