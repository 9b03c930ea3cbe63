#!/bin/bash
# r02 GPU pass 15: ncu of the resident-tile kernel (cp.async halo) and of the Kubo GEMM; GEMM at M = 512 / 514 / 640; host timing of the ordering
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_res -s 8 -c 1 -f -o /tmp/cubic_res2 \
    python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 16 --reps 0 PBK_RES=1 > gpurun_out/r02_ncu_cubic_res2.log 2>&1
ncu -i /tmp/cubic_res2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cubic_res2_raw.csv 2>> gpurun_out/r02_ncu_cubic_res2.log
ncu -i /tmp/cubic_res2.ncu-rep --page source --csv > gpurun_out/r02_ncu_cubic_res2_source.csv 2>> gpurun_out/r02_ncu_cubic_res2.log
tail -2 gpurun_out/r02_ncu_cubic_res2.log
PBK_KUBO_WAVES=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm_kernel -s 1 -c 1 -f -o /tmp/kubo_gemm2 \
    python tools/kubo_bench.py --vectors 1 --reps 0 > gpurun_out/r02_ncu_kubo_gemm2.log 2>&1
ncu -i /tmp/kubo_gemm2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_kubo_gemm2_raw.csv 2>> gpurun_out/r02_ncu_kubo_gemm2.log
ncu -i /tmp/kubo_gemm2.ncu-rep --page source --csv > gpurun_out/r02_ncu_kubo_gemm2_source.csv 2>> gpurun_out/r02_ncu_kubo_gemm2.log
tail -2 gpurun_out/r02_ncu_kubo_gemm2.log
: > gpurun_out/r02_kubo_gemm_shapes.log
for m in 512 514 640 258; do
  echo "# M=$m PBK_KUBO_WAVES=32" >> gpurun_out/r02_kubo_gemm_shapes.log
  PBK_KUBO_WAVES=32 timeout 300 python tools/kubo_bench.py --vectors 1 --reps 1 --moments $m >> gpurun_out/r02_kubo_gemm_shapes.log 2>&1
done
cut -c1-260 gpurun_out/r02_kubo_gemm_shapes.log
PBK_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/r02_bench_timing_v4.json 2> gpurun_out/r02_host_timing_v4.log; echo "bench exit $?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_timing_v4.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e'])"; grep "cluster_order\|set_hamiltonian" gpurun_out/r02_host_timing_v4.log | tail -7
