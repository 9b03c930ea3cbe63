#!/bin/bash
# GPU pass 12: tiles of consecutive rows in the caller's order (identity order) for the cubic lattice, ncu of the two-step kernel.
mkdir -p gpurun_out
timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 128 --reps 1 \
  MB=64 MB=64,PBK_IDENTITY_ORDER=1 MB=64,PBK_IDENTITY_ORDER=1,PBK_TILE=1024 MB=64,PBK_IDENTITY_ORDER=1,PBK_TILE=64 MB=32,PBK_IDENTITY_ORDER=1 MB=128,PBK_IDENTITY_ORDER=1 \
  MB=64,PBK_IDENTITY_ORDER=1,PBK_BULK=6 MB=64,PBK_IDENTITY_ORDER=1,PBK_XS=0 > gpurun_out/sweep_cubic_identity.log 2>&1
timeout 600 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 34 --reps 1 \
  MB=64 MB=64,PBK_IDENTITY_ORDER=1 MB=32,PBK_IDENTITY_ORDER=1 > gpurun_out/sweep_graphene_identity.log 2>&1
for r in 32 64; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheb_pair_bulk -s 4 -c 1 -f -o gpurun_out/pair_r$r \
  python tools/step_sweep.py --workload graphene_200nm_c64_dos --moments 34 --reps 0 PBK_PAIR=1,MB=$r,PBK_PAIR_MINB=2 > gpurun_out/ncu_pair_r$r.log 2>&1
ncu -i gpurun_out/pair_r$r.ncu-rep --page raw --csv > gpurun_out/pair_r${r}_raw.csv 2>/dev/null
ncu -i gpurun_out/pair_r$r.ncu-rep --page source --csv > gpurun_out/pair_r${r}_source.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/sweep_cubic_identity.log gpurun_out/sweep_graphene_identity.log; tail -n 3 gpurun_out/ncu_pair_r32.log
