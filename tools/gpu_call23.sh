#!/bin/bash
# GPU pass 23 (final of round 1): full GPU suite, smoke, bench at HEAD, host-side phase timings.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
PBK_TIMING=1 timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; cat gpurun_out/bench_full.json; grep pbkpm gpurun_out/bench_full.err | tail -n 14
