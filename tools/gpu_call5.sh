#!/bin/bash
# GPU pass 5: spread-LDOS path, configs[2] / configs[3] end to end, split-K wave sweep of the Kubo GEMM.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for w in 4 8 16 32 64; do PBK_KUBO_WAVES=$w timeout 300 python tools/kubo_bench.py --reps 1 2>> gpurun_out/kubo_waves.err | sed "s/^/waves=$w /" >> gpurun_out/kubo_waves.log; done
timeout 900 python tools/config_bench.py sigma > gpurun_out/cfg_sigma.json 2> gpurun_out/cfg_sigma.err
timeout 1200 python tools/config_bench.py ldos > gpurun_out/cfg_ldos.json 2> gpurun_out/cfg_ldos.err
timeout 900 python tools/config_bench.py greens > gpurun_out/cfg_greens.json 2> gpurun_out/cfg_greens.err
tail -n 12 gpurun_out/pytest_gpu.log; cat gpurun_out/kubo_waves.log gpurun_out/cfg_sigma.json gpurun_out/cfg_ldos.json gpurun_out/cfg_greens.json; tail -n 3 gpurun_out/cfg_*.err gpurun_out/kubo_waves.err
