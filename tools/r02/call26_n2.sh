#!/bin/bash
# r02 GPU pass 26 (2 GPUs): the sharded benchmark as the driver launches it; LDOS sites sharded over two ranks
mkdir -p gpurun_out
PBK_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_full_n2_v2.json 2> gpurun_out/r02_bench_full_n2_v2.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_full_n2_v2.json'))
print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['seconds'],d['e2e']['set_model_seconds'],d['clocks'],d['parity']['parity_max_rel'])"
grep "set_hamiltonian\|Error\|error" gpurun_out/r02_bench_full_n2_v2.err | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/ldos_n2.py > gpurun_out/r02_ldos_n2_v2.log 2>&1; tail -2 gpurun_out/r02_ldos_n2_v2.log
