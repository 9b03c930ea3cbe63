#!/bin/bash
# r02 GPU pass 50: ncu --set full of one GROUPED light-cone step (64 sub-systems) of configs[2] LDOS
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:cone_group_step_kernel -s 2200 -c 1 -f -o /tmp/cone_step \
    python bench.py --workload graphene_500nm_c128_ldos --steps 1 --warmup 1 > gpurun_out/r02_ncu_cone_step.log 2>&1
ncu -i /tmp/cone_step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cone_step_raw.csv 2>> gpurun_out/r02_ncu_cone_step.log
tail -1 gpurun_out/r02_ncu_cone_step.log | cut -c1-160
