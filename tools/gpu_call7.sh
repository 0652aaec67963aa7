#!/bin/bash
# 8-GPU call: multi-rank parity of the sequence-parallel path over NCCL (both exchange modes), then bench with the also-array.
out=gpurun_out/r02_call7; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for ex in allgather sendrecv; do
  NCCL_DEBUG=WARN timeout 240 $TR --master-port 29511 tools/seqpar_check.py --seconds 60 --sub-chunks 74 --exchange $ex > $out/seqpar_check_8gpu_$ex.json 2> $out/seqpar_check_$ex.err; echo "seqpar $ex rc=$?" | tee -a $out/summary.txt
  cat $out/seqpar_check_8gpu_$ex.json; tail -3 $out/seqpar_check_$ex.err
done
NCCL_DEBUG=WARN timeout 600 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err; echo "bench8 rc=$?" | tee -a $out/summary.txt
cut -c1-400 $out/bench_8gpu.json; tail -5 $out/bench_8gpu.err
