#!/bin/bash
# GPU pass 6: pybind11 binding, device reconstruction, GEMM warp-layout sweep.
mkdir -p gpurun_out; rm -f gpurun_out/kubo_wn.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for cfg in "2 4" "2 8" "4 4" "4 8" "4 16"; do set -- $cfg; PBK_KUBO_WN=$1 PBK_KUBO_WAVES=$2 timeout 300 python tools/kubo_bench.py --reps 2 2>> gpurun_out/kubo_wn.err | sed "s/^/wn=$1 waves=$2 /" >> gpurun_out/kubo_wn.log; done
PBK_KUBO_WN=4 timeout 300 python tools/kubo_bench.py --reps 1 --moments 1026 --dtype float32 2>> gpurun_out/kubo_wn.err | sed "s/^/wn=4 f32 M=1026 /" >> gpurun_out/kubo_wn.log
PBK_KUBO_WN=4 timeout 300 python tools/kubo_bench.py --reps 1 --size 120 --dtype complex128 2>> gpurun_out/kubo_wn.err | sed "s/^/wn=4 c128 /" >> gpurun_out/kubo_wn.log
timeout 1200 python tools/config_bench.py ldos > gpurun_out/cfg_ldos.json 2> gpurun_out/cfg_ldos.err
timeout 900 python tools/config_bench.py greens > gpurun_out/cfg_greens.json 2> gpurun_out/cfg_greens.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm -c 1 -f -o gpurun_out/kubo_gemm_v3 \
  python tools/kubo_bench.py --size 100 --reps 0 > gpurun_out/ncu_kubo.log 2>&1
tail -n 15 gpurun_out/pytest_gpu.log; cat gpurun_out/kubo_wn.log gpurun_out/cfg_ldos.json gpurun_out/cfg_greens.json; tail -n 3 gpurun_out/cfg_*.err gpurun_out/kubo_wn.err
