#!/bin/bash
# r02 GPU pass 27 (4 GPUs): the sharded benchmark as the driver launches it
mkdir -p gpurun_out
PBK_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r02_bench_full_n4_v2.json 2> gpurun_out/r02_bench_full_n4_v2.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_full_n4_v2.json'))
print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['seconds'],d['e2e']['set_model_seconds'],d['clocks'],d['parity']['parity_max_rel'],d['moment_checksum'])"
grep "set_hamiltonian" gpurun_out/r02_bench_full_n4_v2.err | tail -8; grep -i "error" gpurun_out/r02_bench_full_n4_v2.err | tail -3
