#!/bin/bash
# GPU pass 10: grouped light-cone LDOS, CUDA-graph replay of small recursions, configs[0] / configs[2] numbers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "cone or graph or ldos or spread" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/config_bench.py ldos > gpurun_out/cfg_ldos.json 2> gpurun_out/cfg_ldos.err
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/bench_40nm.json 2> gpurun_out/bench_40nm.err
PBK_GRAPH=0 timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_40nm_nograph.json 2> gpurun_out/bench_40nm_nograph.err
tail -n 12 gpurun_out/pytest_new.log; tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/cfg_ldos.json; tail -n 3 gpurun_out/cfg_ldos.err; cat gpurun_out/bench_40nm.json gpurun_out/bench_40nm_nograph.json; tail -n 3 gpurun_out/bench_40nm*.err
