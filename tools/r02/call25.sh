#!/bin/bash
# r02 GPU pass 25: the headline workload under the power cap (5 s runs): resident-tile kernel on graphene, tile size x vectors per pass of the staged kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "resident" 2>&1 | tail -2
timeout 2400 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 1026 --reps 0 \
  PBK_TILE=256 \
  PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=256,PBK_RES_CTAS=2 PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=384,PBK_RES_CTAS=2 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_CTAS=3 PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=512,PBK_RES_CTAS=3 \
  PBK_RES=2,PBK_RES_ROW=512,PBK_RES_TILE=256,PBK_RES_CTAS=1 \
  MB=32,PBK_TILE=128 MB=64,PBK_TILE=128 MB=64,PBK_TILE=64 MB=32,PBK_TILE=256 MB=16,PBK_TILE=256 MB=16,PBK_TILE=512 \
  > gpurun_out/r02_sweep_headline_powercap_v2.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos\|resident-tile" gpurun_out/r02_sweep_headline_powercap_v2.log | cut -c1-400
