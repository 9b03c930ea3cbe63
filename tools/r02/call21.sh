#!/bin/bash
# r02 GPU pass 21: racecheck of the GEMM with individual hazard records (type, threads); conductivity wall time with cached stacks
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 12 python tools/sanitizer_gemm.py 128 > gpurun_out/r02_racecheck_gemm_hazards.log 2>&1
head -60 gpurun_out/r02_racecheck_gemm_hazards.log | cut -c1-250
PBK_TIMING=1 timeout 600 python bench.py --workload graphene_200nm_f64_conductivity --steps 3 --warmup 1 > gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v4.json 2> gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v4.err
cut -c1-330 gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v4.json; grep "calc_conductivity\|moments_kubo" gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v4.err | tail -10
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "kubo or conductivity" 2>&1 | tail -2
