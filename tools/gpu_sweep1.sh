#!/bin/bash
# Sweep of site ordering / CTA tiling for the step kernel + ncu full captures (natural vs clustered order).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python tools/step_sweep.py --workload graphene_200nm_c64_dos --moments 130 \
  PBK_TILE=-1 PBK_TILE=-1,PBK_TPB=512 PBK_TILE=64 PBK_TILE=128 PBK_TILE=256 PBK_TILE=512 PBK_TILE=1024 PBK_TILE=4096 \
  PBK_TILE=256,PBK_TPB=512 PBK_TILE=512,PBK_TPB=512 PBK_TILE=1024,PBK_TPB=1024 PBK_TILE=256,PBK_BPSM=2 PBK_TILE=256,PBK_BPSM=3 \
  > gpurun_out/sweep_200nm.log 2>&1
timeout 900 python tools/step_sweep.py --workload graphene_200nm_c64_dos --moments 130 --vectors 8 \
  PBK_TILE=-1 PBK_TILE=256 PBK_TILE=1024 PBK_TILE=4096 PBK_TILE=1024,PBK_TPB=512 \
  > gpurun_out/sweep_200nm_r8.log 2>&1
for t in -1 256; do
PBK_TILE=$t timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheb_step -s 20 -c 2 -f -o gpurun_out/step_tile$t \
  python tools/step_sweep.py --workload graphene_200nm_c64_dos --moments 34 --reps 0 PBK_TILE=$t > gpurun_out/ncu_tile$t.log 2>&1
done
timeout 1200 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 PBK_TILE=-1 PBK_TILE=256 PBK_TILE=1024 > gpurun_out/sweep_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_200nm.log gpurun_out/sweep_200nm_r8.log gpurun_out/sweep_full.log
