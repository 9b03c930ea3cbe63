#!/bin/bash
# GPU pass 21: coarse vs fine macro-block pass (throughput + e2e), then the full GPU suite at HEAD.
mkdir -p gpurun_out
timeout 600 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 PBK_COARSE=16 PBK_COARSE=1 PBK_COARSE=16 > gpurun_out/sweep_coarse_full.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/bench_full_coarse.json 2> gpurun_out/bench_full_coarse.err
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/sweep_coarse_full.log gpurun_out/bench_full_coarse.json; tail -n 5 gpurun_out/pytest_gpu.log
