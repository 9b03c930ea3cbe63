#!/bin/bash
# GPU pass 11: host pipeline of set_hamiltonian (parallel copy, eager ordering, pinned staging), Green's timing, full bench + launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
PBK_TIMING=1 timeout 600 python tools/config_bench.py greens > gpurun_out/cfg_greens.json 2> gpurun_out/cfg_greens.err
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/cfg_greens.json; tail -n 5 gpurun_out/cfg_greens.err; cat gpurun_out/bench_full.json; tail -n 3 gpurun_out/bench_full.err; cat gpurun_out/bench_reference.json; tail -n 3 gpurun_out/launches_full.csv
