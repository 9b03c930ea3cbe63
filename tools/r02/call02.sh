#!/bin/bash
# r02 GPU pass 2: full GPU test suite (granule-packed staged kernel, generic ELL widths, multi-orbital model, sweep goldens,
# parity at the BASELINE sizes) and the default bench line with the parity gate and the new CPU baseline.
mkdir -p gpurun_out
rm -f gpurun_out/parity_at_size.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02_pytest_gpu_v1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v1.log
tail -25 gpurun_out/r02_pytest_gpu_v1.log
timeout 900 python bench.py > gpurun_out/r02_bench_full_v1.json 2> gpurun_out/r02_bench_full_v1.err; echo "bench exit $?"
cat gpurun_out/r02_bench_full_v1.json; tail -5 gpurun_out/r02_bench_full_v1.err
