#!/bin/bash
# GPU pass 3: pipelined Kubo-Bastin GEMM (K4) -- parity tests, throughput at configs[3] size, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/kubo_bench.py --reps 2 > gpurun_out/kubo_200nm.json 2> gpurun_out/kubo_200nm.err
timeout 600 python tools/kubo_bench.py --reps 1 --direction xy > gpurun_out/kubo_200nm_xy.json 2>> gpurun_out/kubo_200nm.err
timeout 600 python tools/kubo_bench.py --reps 1 --size 120 --dtype complex128 > gpurun_out/kubo_120nm_c128.json 2>> gpurun_out/kubo_200nm.err
timeout 600 python tools/kubo_bench.py --reps 1 --size 200 --dtype float32 > gpurun_out/kubo_200nm_f32.json 2>> gpurun_out/kubo_200nm.err
timeout 600 python tools/kubo_bench.py --reps 1 --size 200 --moments 1026 --dtype float32 > gpurun_out/kubo_200nm_f32_m1026.json 2>> gpurun_out/kubo_200nm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm -c 1 -f -o gpurun_out/kubo_gemm_v2 \
  python tools/kubo_bench.py --size 100 --reps 0 > gpurun_out/ncu_kubo.log 2>&1
tail -n 12 gpurun_out/pytest_gpu.log; cat gpurun_out/kubo_200nm.json gpurun_out/kubo_200nm_xy.json gpurun_out/kubo_120nm_c128.json gpurun_out/kubo_200nm_f32.json gpurun_out/kubo_200nm_f32_m1026.json; tail -n 5 gpurun_out/kubo_200nm.err; tail -n 3 gpurun_out/ncu_kubo.log
