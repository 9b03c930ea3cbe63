#!/bin/bash
# r02 GPU pass 11 (2 GPUs): the sharded benchmark with the ordering computed on rank 0 and broadcast; LDOS sites sharded over two ranks
mkdir -p gpurun_out
PBK_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_bench_full_n2_v1.json 2> gpurun_out/r02_bench_full_n2_v1.err; echo "bench exit $?"
cat gpurun_out/r02_bench_full_n2_v1.json; grep "pbkpm\|Error\|error" gpurun_out/r02_bench_full_n2_v1.err | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ldos_n2.py > gpurun_out/r02_ldos_n2_v1.log 2>&1; tail -5 gpurun_out/r02_ldos_n2_v1.log
