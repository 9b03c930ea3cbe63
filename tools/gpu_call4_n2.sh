#!/bin/bash
# 2-GPU pass: vectors sharded over two ranks + one ncclAllReduce; N=1 next to it for the checksum and the scaling.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1
timeout 900 python bench.py --workload graphene_200nm_c64_dos --steps 2 --warmup 3 --no-cpu > gpurun_out/n2_bench200_n1.json 2> gpurun_out/n2_bench200_n1.err
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload graphene_200nm_c64_dos --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/n2_bench200_n2.json 2> gpurun_out/n2_bench200_n2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/n2_bench_full_n2.json 2> gpurun_out/n2_bench_full_n2.err
grep -h "NVLS\|NCCL INFO Connected\|Channel 00" gpurun_out/n2_bench200_n2.err | head -8
cat gpurun_out/n2_bench200_n1.json gpurun_out/n2_bench200_n2.json gpurun_out/n2_bench_full_n2.json; tail -n 4 gpurun_out/n2_bench_full_n2.err
