#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 PBK_TILE=256,PBK_PF=4 PBK_TILE=256,PBK_PF=3 PBK_TILE=256,PBK_PF=5 PBK_TILE=256,PBK_PF=6 \
   PBK_TILE=256,PBK_PF=4,PBK_PFMASK=7 PBK_TILE=256,PBK_PF=4,PBK_PFMASK=1 PBK_TILE=256,PBK_PF=8,PBK_PFMASK=7 PBK_TILE=256,PBK_PF=6,PBK_PFMASK=7 \
   PBK_TILE=256,PBK_TPB=2563,PBK_PF=6 PBK_TILE=256,PBK_TPB=2563,PBK_PF=8,PBK_PFMASK=7 PBK_TILE=256,PBK_TPB=2563,PBK_PF=10 PBK_TILE=256,PBK_TPB=2563,PBK_PF=12 PBK_TILE=512,PBK_TPB=2563,PBK_PF=8 PBK_TILE=128,PBK_TPB=2563,PBK_PF=8 > gpurun_out/sweep4_full.log 2>&1
timeout 1500 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --vectors 8 --reps 1 PBK_TILE=-1 PBK_TILE=256 PBK_TILE=256,PBK_PF=4 PBK_TILE=256,PBK_PF=8 PBK_TILE=256,PBK_TPB=2563,PBK_PF=8 PBK_TILE=1024,PBK_PF=4 PBK_TILE=256,PBK_PF=4,PBK_PFMASK=1 > gpurun_out/sweep4_full_r8.log 2>&1
cat gpurun_out/sweep4_full.log gpurun_out/sweep4_full_r8.log
