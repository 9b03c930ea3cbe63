#!/bin/bash
# r02 GPU pass 20: ncu of the k-blocked GEMM; racecheck of the GEMM by tile shape; where the conductivity call spends its wall time; chunked copies A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "resident" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm_kernel -c 1 -f -o /tmp/kubo_gemm3 \
    python tools/kubo_bench.py --vectors 1 --reps 0 > gpurun_out/r02_ncu_kubo_gemm3.log 2>&1
ncu -i /tmp/kubo_gemm3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_kubo_gemm3_raw.csv 2>> gpurun_out/r02_ncu_kubo_gemm3.log
ncu -i /tmp/kubo_gemm3.ncu-rep --page source --csv > gpurun_out/r02_ncu_kubo_gemm3_source.csv 2>> gpurun_out/r02_ncu_kubo_gemm3.log
tail -2 gpurun_out/r02_ncu_kubo_gemm3.log
: > gpurun_out/r02_racecheck_gemm.log
for m in 128 134 18 200; do
  echo "##### M=$m" >> gpurun_out/r02_racecheck_gemm.log
  timeout 600 compute-sanitizer --tool racecheck --print-limit 3 python tools/sanitizer_gemm.py $m >> gpurun_out/r02_racecheck_gemm.log 2>&1
done
grep "#####\|RACECHECK SUMMARY\|gemm launches ok" gpurun_out/r02_racecheck_gemm.log
PBK_TIMING=1 timeout 600 python bench.py --workload graphene_200nm_f64_conductivity --steps 2 --warmup 1 > gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v3.json 2> gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v3.err
cut -c1-400 gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v3.json; grep "calc_conductivity\|moments_kubo" gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v3.err | tail -8
: > gpurun_out/r02_kubo_gemm_chunk.log
for cfg in "PBK_KUBO_CHUNK=32 --moments 514" "PBK_KUBO_CHUNK=200 --moments 514" "PBK_KUBO_CHUNK=32 --moments 520" "PBK_KUBO_CHUNK=200 --moments 520" "PBK_KUBO_CHUNK=16 --moments 520" "PBK_KUBO_CHUNK=32 --moments 600" "PBK_KUBO_CHUNK=200 --moments 600" "PBK_KUBO_CHUNK=32 --moments 512" ; do
  echo "# $cfg" >> gpurun_out/r02_kubo_gemm_chunk.log
  env ${cfg%% *} timeout 300 python tools/kubo_bench.py --reps 1 --vectors 2 ${cfg#* } >> gpurun_out/r02_kubo_gemm_chunk.log 2>&1
done
cut -c1-120 gpurun_out/r02_kubo_gemm_chunk.log
