#!/bin/bash
# 2-GPU pass at HEAD (v6 engine) for the scaling table, plus the GPU suite on the final tree.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/n2_bench_full_v6.json 2> gpurun_out/n2_bench_full_v6.err
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
cat gpurun_out/n2_bench_full_v6.json; tail -n 4 gpurun_out/pytest_gpu_final.log
