#!/bin/bash
# r02 GPU pass 1: ncu --set full of cheb_step_bulk<float,4,7,...> on configs[4] (cubic 256^3) at 64 and 16 vectors per pass,
# plus CTAs-per-SM / pipeline-depth sweeps of the same workload (is it L1 capacity, L2 bandwidth or latency?).
mkdir -p gpurun_out
for R in 64 16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_bulk -s 8 -c 1 -f -o /tmp/cubic_r$R \
    python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors $R --reps 0 MB=$R > gpurun_out/r02_ncu_cubic_r$R.log 2>&1
  ncu -i /tmp/cubic_r$R.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cubic_r${R}_raw.csv 2>> gpurun_out/r02_ncu_cubic_r$R.log
  ncu -i /tmp/cubic_r$R.ncu-rep --page source --csv > gpurun_out/r02_ncu_cubic_r${R}_source.csv 2>> gpurun_out/r02_ncu_cubic_r$R.log
  ls -la /tmp/cubic_r$R.ncu-rep >> gpurun_out/r02_ncu_cubic_r$R.log
done
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  MB=64 MB=64,PBK_BPSM=1 MB=64,PBK_BPSM=2 MB=64,PBK_BPSM=3 MB=64,PBK_BULK=2 MB=64,PBK_BULK=2,PBK_BPSM=2 MB=64,PBK_XS=0 MB=64,PBK_XS=0,PBK_BPSM=2 \
  MB=16 MB=16,PBK_BPSM=2 MB=16,PBK_TILE=1024 MB=16,PBK_TILE=1024,PBK_BPSM=2 MB=32,PBK_BPSM=2 > gpurun_out/r02_sweep_cubic_bpsm.log 2>&1
cat gpurun_out/r02_sweep_cubic_bpsm.log
tail -3 gpurun_out/r02_ncu_cubic_r64.log gpurun_out/r02_ncu_cubic_r16.log
