#!/bin/bash
# 1-GPU evidence call on the round's final sources: default bench (cfg2 + also-array), scan DRAM traffic (ncu), ncu --set full
# of the scan and of the streaming push kernel, ncu metric pass of the wide GEMMs, ncu launch list of the bench command.
out=gpurun_out/r02_call11; mkdir -p $out
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" | tee -a $out/summary.txt
cut -c1-600 $out/bench_default.json; tail -3 $out/bench_default.err
timeout 900 python tools/scan_traffic.py --out $out/scan_traffic.json > $out/traffic.log 2>&1; tail -4 $out/traffic.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel_pair -s 3 -c 1 -o $out/scan_pair_S_fp32 -f python tools/scan_bench.py --hparams S --batch 32 --L 3999 --mode fp32 --iters 2 > $out/ncu_scan.log 2>&1; echo "ncu scan rc=$?" | tee -a $out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_push_kernel -s 20 -c 1 -o $out/stream_push_S_f20 -f python tools/stream_push_timeline.py --frames 20 --steps 30 > $out/ncu_stream.log 2>&1; echo "ncu stream rc=$?" | tee -a $out/summary.txt
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"
for g in in_proj out_proj x_proj; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 3 -c 1 --csv --log-file $out/gemm_${g}_S_fp32.csv python tools/gemm_bench.py --only $g --iters 3 > /dev/null 2>&1; echo "ncu gemm $g rc=$?" | tee -a $out/summary.txt
  timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 3 -c 1 --csv --log-file $out/gemm_${g}_L_bf16.csv python tools/gemm_bench.py --only $g --iters 3 --hparams L --batch 64 --mode bf16 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 250 --csv --log-file $out/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $out/summary.txt
ls -la $out
