#!/bin/bash
# r02 GPU pass 3: full GPU test suite (device-built layouts, granule-packed staged kernel, generic ELL widths, multi-orbital
# model, sweep goldens, parity at the BASELINE sizes), host-pipeline timing at 38 M sites, compute-sanitizer.
mkdir -p gpurun_out
rm -f gpurun_out/parity_at_size.jsonl
timeout 1800 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02_pytest_gpu_v2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v2.log
tail -40 gpurun_out/r02_pytest_gpu_v2.log
PBK_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/r02_bench_timing_v2.json 2> gpurun_out/r02_host_timing_v2.log; echo "bench exit $?"
cat gpurun_out/r02_bench_timing_v2.json; grep pbkpm gpurun_out/r02_host_timing_v2.log | tail -40
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitizer_cases.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log
