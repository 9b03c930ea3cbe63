#!/bin/bash
# r02 GPU pass 19 (checkpoint): full GPU test suite, smoke, every bench workload, ncu of the new GEMM and of the resident-tile kernel, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_v4.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_v4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v1.log 2>&1; tail -2 gpurun_out/r02_smoke_v1.log
for wl in graphene_200nm_f64_conductivity graphene_500nm_c128_ldos graphene_500nm_c128_greens; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 > gpurun_out/r02_bench_${wl}_v2.json 2> gpurun_out/r02_bench_${wl}_v2.err; cut -c1-1400 gpurun_out/r02_bench_${wl}_v2.json; tail -2 gpurun_out/r02_bench_${wl}_v2.err
done
timeout 600 python bench.py --workload cubic_256_f32_dos --steps 2 --warmup 3 > gpurun_out/r02_bench_cubic_v1.json 2> gpurun_out/r02_bench_cubic_v1.err; cut -c1-2500 gpurun_out/r02_bench_cubic_v1.json; tail -2 gpurun_out/r02_bench_cubic_v1.err
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v2.json 2> gpurun_out/r02_bench_40nm_v2.err; cut -c1-1200 gpurun_out/r02_bench_40nm_v2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm_kernel -s 1 -c 1 -f -o /tmp/kubo_gemm3 \
    python tools/kubo_bench.py --vectors 1 --reps 0 > gpurun_out/r02_ncu_kubo_gemm3.log 2>&1
ncu -i /tmp/kubo_gemm3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_kubo_gemm3_raw.csv 2>> gpurun_out/r02_ncu_kubo_gemm3.log
ncu -i /tmp/kubo_gemm3.ncu-rep --page source --csv > gpurun_out/r02_ncu_kubo_gemm3_source.csv 2>> gpurun_out/r02_ncu_kubo_gemm3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_res -s 8 -c 1 -f -o /tmp/cubic_res3 \
    python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 16 --reps 0 PBK_RES=1 > gpurun_out/r02_ncu_cubic_res3.log 2>&1
ncu -i /tmp/cubic_res3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cubic_res3_raw.csv 2>> gpurun_out/r02_ncu_cubic_res3.log
ncu -i /tmp/cubic_res3.ncu-rep --page source --csv > gpurun_out/r02_ncu_cubic_res3_source.csv 2>> gpurun_out/r02_ncu_cubic_res3.log
tail -2 gpurun_out/r02_ncu_kubo_gemm3.log gpurun_out/r02_ncu_cubic_res3.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_cases.py > gpurun_out/r02_sanitizer_${tool}_v2.log 2>&1; tail -4 gpurun_out/r02_sanitizer_${tool}_v2.log
done
