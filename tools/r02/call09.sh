#!/bin/bash
# r02 GPU pass 9: persistent kernel, double-buffered resident tiles: tests, cubic sweep, configs[0] bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -12
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_BUFS=1,PBK_RES_CTAS=3,PBK_RES_TILE=384 PBK_RES=1,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_TILE=448 PBK_RES=1,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=256 PBK_RES=1,PBK_RES_TILE=256,PBK_RES_CTAS=3 \
  > gpurun_out/r02_sweep_cubic_res_v4.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian" gpurun_out/r02_sweep_cubic_res_v4.log | cut -c1-330
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v1.json 2> gpurun_out/r02_bench_40nm_v1.err; cat gpurun_out/r02_bench_40nm_v1.json; tail -3 gpurun_out/r02_bench_40nm_v1.err
PBK_PERSIST=0 timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_40nm_v1_graph.json 2>/dev/null; cut -c1-400 gpurun_out/r02_bench_40nm_v1_graph.json
