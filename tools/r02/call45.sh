#!/bin/bash
# r02 GPU pass 45: GEMM edge tiles with more, smaller stages (parity by tile shape, rates by M), racecheck of an edge-tile case
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -k "kubo or conductivity" 2>&1 | tail -2
: > gpurun_out/r02_kubo_gemm_shapes_v4.log
for cfg in "--moments 514 --vectors 4" "--moments 520 --vectors 2" "--moments 544 --vectors 2" "--moments 600 --vectors 2" "--moments 258 --vectors 4" "--moments 200 --vectors 4" "--moments 1398 --vectors 1"; do
  echo "# $cfg" >> gpurun_out/r02_kubo_gemm_shapes_v4.log
  timeout 300 python tools/kubo_bench.py --reps 1 $cfg >> gpurun_out/r02_kubo_gemm_shapes_v4.log 2>&1
done
cut -c1-120 gpurun_out/r02_kubo_gemm_shapes_v4.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 2 python tools/sanitizer_gemm.py 200 2>&1 | grep "RACECHECK\|launches ok"
timeout 300 compute-sanitizer --tool memcheck --print-limit 2 python tools/sanitizer_gemm.py 136 2>&1 | grep "ERROR SUMMARY\|launches ok"
