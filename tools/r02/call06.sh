#!/bin/bash
# r02 GPU pass 6: stage hand-back variants of the staged kernel at 38 M sites; sweep goldens; resident-tile kernel smoke + sweep on the cubic lattice
mkdir -p gpurun_out
timeout 1200 python tools/r02/diag_r4c.py > gpurun_out/r02_diag_r4c.log 2>&1; echo "diag exit $?"
cat gpurun_out/r02_diag_r4c.log
timeout 300 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -5
PBK_RES=2 PBK_RES_TILE=128 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dos_moments or curves" 2>&1 | tail -8
PBK_TIMING=1 timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=0,MB=64 PBK_RES=0,MB=16 PBK_RES=1 PBK_RES=1,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_TILE=1024,PBK_RES_CTAS=2,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_TILE=256,PBK_RES_CTAS=4 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_CTAS=2 \
  > gpurun_out/r02_sweep_cubic_res_v1.log 2>&1
cat gpurun_out/r02_sweep_cubic_res_v1.log
