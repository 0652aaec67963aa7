#!/bin/bash
# final build: ncu launch list of `bench.py --steps 2 --warmup 1 --no-also` (after the same command exits 0 without ncu), stream bench lines
out=gpurun_out/r02_call30; mkdir -p $out
timeout 600 python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline > $out/b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_cfg2_final2.csv python bench.py --steps 2 --warmup 1 --no-also --no-cpu-baseline > $out/ncu.log 2>&1
echo "ncu rc=$?"; wc -l $out/launches_cfg2_final2.csv
timeout 300 python bench.py --workload stream --causal --batch 1 --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_stream_b1.json 2> $out/stream_b1.err; tail -c 400 $out/bench_stream_b1.json
timeout 300 python bench.py --workload stream --causal --batch 32 --steps 30 --warmup 5 --no-cpu-baseline > $out/bench_stream_b32.json 2> $out/stream_b32.err; tail -c 400 $out/bench_stream_b32.json
