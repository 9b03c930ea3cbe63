#!/bin/bash
# 2-GPU pass at HEAD: DOS sharded over two ranks (+ reference arm under torchrun), LDOS sites sharded over two ranks.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/n2_bench_full_n2.json 2> gpurun_out/n2_bench_full_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/n2_bench_reference.json 2> gpurun_out/n2_bench_reference.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/ldos_n2.py > gpurun_out/n2_ldos.log 2> gpurun_out/n2_ldos.err
cat gpurun_out/n2_bench_full_n2.json; tail -n 3 gpurun_out/n2_bench_full_n2.err; cat gpurun_out/n2_bench_reference.json | cut -c1-300; cat gpurun_out/n2_ldos.log; tail -n 5 gpurun_out/n2_ldos.err
