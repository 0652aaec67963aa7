#!/bin/bash
out=gpurun_out/r02_call20; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_stream_fused.py tests/test_gpu_causal.py -x -q 2>&1 | tail -6 | tee $out/tests.log
timeout 600 python tools/stream_grid.py | tee $out/stream_grid_S_f20.jsonl
timeout 300 python tools/stream_grid.py --hparams L --frames 16 | tee $out/stream_grid_L_f16.jsonl
