#!/bin/bash
# 2-GPU call: the GPU suite (1 GPU), then the multi-rank paths over NCCL: seqpar_check in both exchange modes, bench with the also-array.
out=gpurun_out/r02_call4; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
grep -E "bf16|passed|failed|Error|FAILED" $out/pytest_gpu.log | tail -25
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ex in allgather sendrecv; do
  NCCL_DEBUG=WARN timeout 300 $TR --master-port 29511 tools/seqpar_check.py --seconds 20 --sub-chunks 32 --exchange $ex > $out/seqpar_check_2gpu_$ex.json 2> $out/seqpar_check_$ex.err; echo "seqpar $ex rc=$?" | tee -a $out/summary.txt
  cat $out/seqpar_check_2gpu_$ex.json; tail -3 $out/seqpar_check_$ex.err
done
NCCL_DEBUG=WARN timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench2 rc=$?" | tee -a $out/summary.txt
cut -c1-300 $out/bench_2gpu.json; tail -5 $out/bench_2gpu.err
