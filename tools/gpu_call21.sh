#!/bin/bash
# full GPU suite + default bench line on the current build
out=gpurun_out/r02_call21; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/tests.log
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 600 $out/bench_default.json
