#!/bin/bash
# Round-1 GPU check: parity tests, smoke, bench (reduced + full), ncu launch list. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload graphene_200nm_c64_dos --steps 3 --warmup 3 > gpurun_out/bench_200nm.json 2> gpurun_out/bench_200nm.err
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_200nm.csv python bench.py --workload graphene_200nm_c64_dos --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_200nm.json; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
