#!/bin/bash
# r02 GPU pass 46: ncu of the GEMM at M = 520 (a remainder tile row of 8 rows): what do the edge CTAs cost?
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm_kernel -c 1 -f -o /tmp/kubo_gemm520 \
    python tools/kubo_bench.py --vectors 1 --reps 0 --moments 520 > gpurun_out/r02_ncu_kubo_gemm520.log 2>&1
ncu -i /tmp/kubo_gemm520.ncu-rep --page raw --csv > gpurun_out/r02_ncu_kubo_gemm520_raw.csv 2>> gpurun_out/r02_ncu_kubo_gemm520.log
tail -2 gpurun_out/r02_ncu_kubo_gemm520.log | cut -c1-200
