#!/bin/bash
out=gpurun_out/r02_call12; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_xproj" 2>&1 | tail -15 | tee $out/tests_convx.log
timeout 120 python tools/convx_bench.py | tee $out/convx_S_fp32.json
timeout 120 python tools/convx_bench.py --hparams L --batch 64 --mode bf16 | tee $out/convx_L_bf16.json
timeout 120 python tools/convx_bench.py --mode bf16 | tee $out/convx_S_bf16.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-also --no-cpu-baseline > $out/bench_cfg2.json 2> $out/bench_cfg2.err; cut -c1-300 $out/bench_cfg2.json; tail -3 $out/bench_cfg2.err
python -c "
import json; d=json.loads(open('$out/bench_cfg2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['traffic'], d['clocks'])"
