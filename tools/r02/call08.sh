#!/bin/bash
# r02 GPU pass 8: resident-tile kernel tests + tile sweep + ncu capture on the cubic lattice; batched Kubo-Bastin and light-cone Green's
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -x -k "config2 or config3 or config4" 2>&1 | tail -6
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_TILE=384 PBK_RES=1,PBK_RES_TILE=448 PBK_RES=1,PBK_RES_TILE=640,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_TILE=768,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_TILE=384,PBK_RES_STAGES=3 \
  > gpurun_out/r02_sweep_cubic_res_v3.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian" gpurun_out/r02_sweep_cubic_res_v3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_res -s 8 -c 1 -f -o /tmp/cubic_res \
    python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 16 --reps 0 PBK_RES=1 > gpurun_out/r02_ncu_cubic_res.log 2>&1
ncu -i /tmp/cubic_res.ncu-rep --page raw --csv > gpurun_out/r02_ncu_cubic_res_raw.csv 2>> gpurun_out/r02_ncu_cubic_res.log
ncu -i /tmp/cubic_res.ncu-rep --page source --csv > gpurun_out/r02_ncu_cubic_res_source.csv 2>> gpurun_out/r02_ncu_cubic_res.log
tail -3 gpurun_out/r02_ncu_cubic_res.log
timeout 600 python tools/config_bench.py sigma > gpurun_out/r02_cfg3_sigma_v1.json 2> gpurun_out/r02_cfg3_sigma_v1.err; cat gpurun_out/r02_cfg3_sigma_v1.json; tail -3 gpurun_out/r02_cfg3_sigma_v1.err
timeout 600 python tools/config_bench.py greens > gpurun_out/r02_cfg2_greens_v1.json 2> gpurun_out/r02_cfg2_greens_v1.err; cat gpurun_out/r02_cfg2_greens_v1.json; tail -3 gpurun_out/r02_cfg2_greens_v1.err
