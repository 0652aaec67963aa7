#!/bin/bash
out=gpurun_out/r02_call19; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_stream_fused.py tests/test_gpu_causal.py -x -q 2>&1 | tail -12 | tee $out/tests.log
timeout 600 python tools/stream_grid.py | tee $out/stream_grid_S_f20.jsonl
timeout 600 python tools/stream_grid.py --frames 2 | tee $out/stream_grid_S_f2.jsonl
