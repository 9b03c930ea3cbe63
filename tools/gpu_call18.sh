#!/bin/bash
# GPU pass 18: small tiles under the two-level ordering (L1-resident tile working set).
mkdir -p gpurun_out
timeout 900 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 \
  PBK_MACRO=256 PBK_MACRO=1024,PBK_TILE=64 PBK_MACRO=2048,PBK_TILE=32 PBK_MACRO=1024,PBK_TILE=64,PBK_BULK=6 PBK_MACRO=512,PBK_TILE=128,PBK_BULK=6 \
  PBK_MACRO=1024,PBK_TILE=64,MB=32 PBK_MACRO=512,PBK_TILE=128,MB=32 PBK_MACRO=256,MB=32 > gpurun_out/sweep_macro3_full.log 2>&1
cat gpurun_out/sweep_macro3_full.log
