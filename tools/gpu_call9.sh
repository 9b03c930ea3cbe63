#!/bin/bash
# GPU pass 9: light-cone sub-system LDOS (moments_ldos_cones), full GPU suite, configs[2] LDOS, vectors-per-pass checks.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "cone or pair or ldos" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/config_bench.py ldos > gpurun_out/cfg_ldos.json 2> gpurun_out/cfg_ldos.err
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --max-batch 32 > gpurun_out/bench_full_mb32.json 2> gpurun_out/bench_full_mb32.err
timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 128 --reps 1 \
  MB=64 MB=32 MB=16 MB=128 MB=128,PBK_TILE=512 MB=32,PBK_TILE=1024 > gpurun_out/sweep_cubic_mb.log 2>&1
tail -n 12 gpurun_out/pytest_new.log; tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/cfg_ldos.json; tail -n 3 gpurun_out/cfg_ldos.err; cat gpurun_out/bench_full_mb32.json; cat gpurun_out/sweep_cubic_mb.log
