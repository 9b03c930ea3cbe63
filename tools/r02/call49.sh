#!/bin/bash
# r02 GPU pass 49 (last): GPU suite + smoke on the shipped binary; ncu --set full of one light-cone group step of configs[2] LDOS
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_v9.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_v9.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v4.log 2>&1; tail -1 gpurun_out/r02_smoke_v4.log | cut -c1-120
timeout 600 ncu --set full --clock-control none -k regex:cone_group_step_kernel -s 500 -c 1 -f -o /tmp/cone_step \
    python bench.py --workload graphene_500nm_c128_ldos --steps 1 --warmup 1 > gpurun_out/r02_ncu_cone_step.log 2>&1
ncu -i /tmp/cone_step.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cone_step_raw.csv 2>> gpurun_out/r02_ncu_cone_step.log
tail -1 gpurun_out/r02_ncu_cone_step.log | cut -c1-160
