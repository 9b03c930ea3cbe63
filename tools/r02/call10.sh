#!/bin/bash
# r02 GPU pass 10: resident-tile kernel (3 CTAs/SM, prefetched tile descriptors) sweep; test logs; new bench workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02_pytest_props_v4.log 2>&1; tail -12 gpurun_out/r02_pytest_props_v4.log
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_TILE=512 PBK_RES=1,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_STAGES=3,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_CTAS=4,PBK_RES_TILE=256 \
  > gpurun_out/r02_sweep_cubic_res_v5.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian" gpurun_out/r02_sweep_cubic_res_v5.log | cut -c1-330
for wl in graphene_200nm_f64_conductivity graphene_500nm_c128_ldos graphene_500nm_c128_greens; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 > gpurun_out/r02_bench_${wl}_v1.json 2> gpurun_out/r02_bench_${wl}_v1.err; cut -c1-900 gpurun_out/r02_bench_${wl}_v1.json; tail -2 gpurun_out/r02_bench_${wl}_v1.err
done
