#!/bin/bash
# r02 GPU pass 22: racecheck of the GEMM with per-lane arrivals; GEMM rate check; conductivity step times
mkdir -p gpurun_out
: > gpurun_out/r02_racecheck_gemm_v2.log
for m in 128 134 200; do
  echo "##### M=$m" >> gpurun_out/r02_racecheck_gemm_v2.log
  timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python tools/sanitizer_gemm.py $m >> gpurun_out/r02_racecheck_gemm_v2.log 2>&1
done
grep "#####\|RACECHECK SUMMARY\|gemm launches ok" gpurun_out/r02_racecheck_gemm_v2.log
for cfg in "--moments 514 --vectors 4" "--moments 512 --vectors 2"; do timeout 300 python tools/kubo_bench.py --reps 2 $cfg 2>&1 | cut -c1-150; done
timeout 600 python bench.py --workload graphene_200nm_f64_conductivity --steps 3 --warmup 1 > gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v5.json 2> gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v5.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v5.json'));print(d['value'],d['ms_per_step'],d['config']['step_seconds'],d['roofline']['frac'],d['roofline']['gemm_ms_per_call'],d['roofline']['recursion_ms_per_call'])"
