#!/bin/bash
# r02 GPU pass 24: the headline step under the power cap (5 s runs): f32 tile-partial sums, resident-tile kernel on graphene, ring variants
mkdir -p gpurun_out
timeout 1500 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 1026 --reps 0 \
  PBK_TILE=256 PBK_FSUMS=1 \
  PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=256,PBK_RES_CTAS=2 PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=384,PBK_RES_CTAS=2,PBK_RES_STAGES=2 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_CTAS=3 PBK_RES=2,PBK_RES_ROW=512,PBK_RES_TILE=256,PBK_RES_CTAS=1 \
  PBK_XS=0 PBK_BULK=6 PBK_FSUMS=1,PBK_XS=0 PBK_TILE=256 \
  > gpurun_out/r02_sweep_headline_powercap_v1.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos\|resident-tile" gpurun_out/r02_sweep_headline_powercap_v1.log | cut -c1-420
