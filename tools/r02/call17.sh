#!/bin/bash
# r02 GPU pass 17: Kubo GEMM on k-blocked stacks (parity + shapes); resident-tile kernel A/B with a clean environment per config
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "kubo or conductivity or multi_orbital" > gpurun_out/r02_pytest_kubo_v1.log 2>&1; tail -5 gpurun_out/r02_pytest_kubo_v1.log
: > gpurun_out/r02_kubo_gemm_shapes_v2.log
for cfg in "--moments 514 --vectors 1" "--moments 512 --vectors 1" "--moments 514 --vectors 4" "--moments 258 --vectors 4" "--moments 514 --vectors 2 --dtype complex128" "--moments 514 --vectors 4 --dtype float32"; do
  echo "# $cfg" >> gpurun_out/r02_kubo_gemm_shapes_v2.log
  timeout 300 python tools/kubo_bench.py --reps 1 $cfg >> gpurun_out/r02_kubo_gemm_shapes_v2.log 2>&1
done
for w in 8 32; do echo "# PBK_KUBO_WAVES=$w --moments 514 --vectors 4" >> gpurun_out/r02_kubo_gemm_shapes_v2.log; PBK_KUBO_WAVES=$w timeout 300 python tools/kubo_bench.py --reps 1 --moments 514 --vectors 4 >> gpurun_out/r02_kubo_gemm_shapes_v2.log 2>&1; done
cut -c1-250 gpurun_out/r02_kubo_gemm_shapes_v2.log
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_L2PF=0 PBK_RES=1,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_STAGES=3,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_TILE=448 \
  PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_ROW=32,PBK_RES_TILE=512 PBK_RES=1,PBK_RES_ROW=32,PBK_RES_TILE=768,PBK_RES_CTAS=4 \
  > gpurun_out/r02_sweep_cubic_res_v10.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos" gpurun_out/r02_sweep_cubic_res_v10.log | cut -c1-250
