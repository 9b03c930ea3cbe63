#!/bin/bash
# GPU pass 17: re-sweep of the step kernel's knobs under the two-level ordering; per-GPU share of the 8-GPU sharding.
mkdir -p gpurun_out
timeout 900 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 \
  PBK_MACRO=256 PBK_MACRO=256,PBK_TILE=128 PBK_MACRO=512,PBK_TILE=128 PBK_MACRO=128,PBK_TILE=512 PBK_MACRO=256,PBK_BULK=3 PBK_MACRO=256,PBK_BULK=6 \
  PBK_MACRO=256,PBK_XS=0 PBK_MACRO=256,PBK_BPSM=3 PBK_MACRO=4096 PBK_MACRO=256,MB=8 PBK_MACRO=0,MB=8 PBK_MACRO=256,MB=16 > gpurun_out/sweep_macro2_full.log 2>&1
cat gpurun_out/sweep_macro2_full.log
