#!/bin/bash
# r02 GPU pass 7: full GPU suite with the proxy-fence fix and the resident-tile kernel; cubic sweep of the resident kernel
mkdir -p gpurun_out
rm -f gpurun_out/parity_at_size.jsonl
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu_v3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v3.log
tail -30 gpurun_out/r02_pytest_gpu_v3.log
PBK_TIMING=1 timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=0,MB=64 PBK_RES=1 PBK_RES=1,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_TILE=1024,PBK_RES_CTAS=2,PBK_RES_STAGES=3 PBK_RES=1,PBK_RES_TILE=256,PBK_RES_CTAS=4 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_CTAS=3 \
  > gpurun_out/r02_sweep_cubic_res_v2.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian" gpurun_out/r02_sweep_cubic_res_v2.log
timeout 600 python tools/r02/diag_r4c.py > gpurun_out/r02_diag_r4c_fixed.log 2>&1; grep '"env": {}' gpurun_out/r02_diag_r4c_fixed.log | cut -c1-200
