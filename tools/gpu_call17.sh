#!/bin/bash
out=gpurun_out/r02_call17; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stream_fused.py tests/test_gpu_causal.py -x -q 2>&1 | tail -8 | tee $out/tests.log
for t in 1 8 20; do timeout 200 python tools/stack_step_latency.py --tokens $t | tee -a $out/stack_step_latency.jsonl; done
timeout 200 python tools/stack_step_latency.py --tokens 1 --batch 8 | tee -a $out/stack_step_latency.jsonl
timeout 300 python bench.py --workload stream --batch 1 --steps 200 --warmup 10 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream b1', d['ms_per_step'])"
