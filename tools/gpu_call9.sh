#!/bin/bash
out=gpurun_out/r02_call9; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_stream_fused.py -x -q 2>&1 | tail -5 | tee $out/tests_stream_fused.log
for fr in 20 2 32; do timeout 120 python tools/stream_push_timeline.py --frames $fr | tee $out/timeline_S_b1_f$fr.json; done
timeout 120 python tools/stream_push_timeline.py --frames 20 --batch 32 | tee $out/timeline_S_b32_f20.json
