#!/bin/bash
# GPU pass 7: re-check HEAD (tests + smoke), ncu capture of the Kubo-Bastin GEMM (tensor pipe), per-GPU share of the 8-GPU sharding.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm -c 1 -f -o gpurun_out/kubo_gemm_r01 \
  python tools/kubo_bench.py --reps 0 > gpurun_out/ncu_kubo.log 2>&1
ncu -i gpurun_out/kubo_gemm_r01.ncu-rep --page raw --csv > gpurun_out/kubo_gemm_r01_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
for mb in 8 16; do timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --max-batch $mb > gpurun_out/bench_full_mb$mb.json 2> gpurun_out/bench_full_mb$mb.err; done
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -n 3 gpurun_out/ncu_kubo.log; cat gpurun_out/bench_full_mb*.json; tail -n 3 gpurun_out/*.err
