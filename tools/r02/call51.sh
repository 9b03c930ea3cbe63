#!/bin/bash
# r02 GPU pass 51: the headline step under the power cap with fewer resident CTAs per SM / a shallower ring
mkdir -p gpurun_out
timeout 600 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 1026 --reps 0 PBK_TILE=256 PBK_BPSM=3 PBK_BPSM=2 PBK_BULK=3 PBK_BPSM=3,PBK_BULK=6 > gpurun_out/r02_sweep_headline_powercap_v4.log 2>&1
grep -v "pbkpm" gpurun_out/r02_sweep_headline_powercap_v4.log | cut -c1-330
