#!/bin/bash
# r02 GPU pass 48: ncu launch lists of the conductivity and the cubic workloads (kernel shares of the step)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_conductivity_v1.csv python tools/kubo_bench.py --vectors 4 --reps 0 > gpurun_out/r02_launches_conductivity_v1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cubic_v1.csv python tools/step_sweep.py --workload cubic_256_f32_dos --moments 258 --vectors 32 --reps 0 PBK_RES=1 > gpurun_out/r02_launches_cubic_v1.log 2>&1
wc -l gpurun_out/r02_launches_conductivity_v1.csv gpurun_out/r02_launches_cubic_v1.csv
