#!/bin/bash
# r02 GPU pass 16: resident-tile kernel with the descriptor cache and L2 prefetch of the next tile's streams
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "resident" > gpurun_out/r02_pytest_res_v7.log 2>&1; tail -4 gpurun_out/r02_pytest_res_v7.log
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_L2PF=0 PBK_RES=1,PBK_RES_TILE=512 PBK_RES=1,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_STAGES=3,PBK_RES_TILE=320 PBK_RES=1,PBK_RES_STAGES=3,PBK_RES_TILE=256 \
  PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_CTAS=4,PBK_RES_TILE=256 PBK_RES=1,PBK_RES_CTAS=2,PBK_RES_TILE=512,PBK_RES_STAGES=4 \
  > gpurun_out/r02_sweep_cubic_res_v9.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos" gpurun_out/r02_sweep_cubic_res_v9.log | cut -c1-250
