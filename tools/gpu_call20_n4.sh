#!/bin/bash
# 4-GPU pass: configs[1] exactly as named (64 vectors sharded over 4 ranks, one ncclAllReduce).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n4_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 2 --warmup 2 > gpurun_out/n4_bench_full.json 2> gpurun_out/n4_bench_full.err
cat gpurun_out/n4_bench_full.json; tail -n 4 gpurun_out/n4_bench_full.err; wc -l gpurun_out/n4_gpus.txt
