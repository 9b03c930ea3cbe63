#!/bin/bash
# r02 GPU pass 12: where does the e2e time outside the moments phase go; double-buffered resident tiles at 3 CTAs/SM; property tests
mkdir -p gpurun_out
PBK_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/r02_bench_timing_v3.json 2> gpurun_out/r02_host_timing_v3.log; echo "bench exit $?"
python -c "import json;d=json.load(open('gpurun_out/r02_bench_timing_v3.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e'])"; grep "calc_dos\|moments_dos\|set_hamiltonian" gpurun_out/r02_host_timing_v3.log | tail -8
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=256 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=384,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=128,PBK_RES_CTAS=4 \
  > gpurun_out/r02_sweep_cubic_res_v6.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos" gpurun_out/r02_sweep_cubic_res_v6.log | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02_pytest_props_v5.log 2>&1; tail -6 gpurun_out/r02_pytest_props_v5.log
