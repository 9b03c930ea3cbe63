#!/bin/bash
# r02 GPU pass 30: resident-tile kernel on the headline workload under the power cap, double-buffered tiles and deeper rings
mkdir -p gpurun_out
timeout 2400 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 1026 --reps 0 \
  PBK_TILE=256 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_BUFS=2,PBK_RES_CTAS=2 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_BUFS=2,PBK_RES_CTAS=2,PBK_RES_STAGES=4 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=512,PBK_RES_CTAS=2,PBK_RES_STAGES=4 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=768,PBK_RES_CTAS=2,PBK_RES_STAGES=4 \
  PBK_RES=2,PBK_RES_ROW=128,PBK_RES_TILE=384,PBK_RES_CTAS=4,PBK_RES_STAGES=2 \
  PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=128,PBK_RES_BUFS=2,PBK_RES_CTAS=2 \
  PBK_RES=2,PBK_RES_ROW=256,PBK_RES_TILE=256,PBK_RES_CTAS=2,PBK_RES_STAGES=4 \
  PBK_RES=2,PBK_RES_ROW=64,PBK_RES_TILE=1024,PBK_RES_CTAS=3 \
  > gpurun_out/r02_sweep_headline_powercap_v3.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos\|resident-tile" gpurun_out/r02_sweep_headline_powercap_v3.log | cut -c1-330
