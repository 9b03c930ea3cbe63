#!/bin/bash
# r02 GPU pass 13: K4 GEMM stage width / warp layout / split-K waves; resident kernel with 128-byte rows; Kubo parity
mkdir -p gpurun_out
: > gpurun_out/r02_kubo_gemm_sweep.log
for cfg in "PBK_KUBO_ROW=128 PBK_KUBO_WN=4" "PBK_KUBO_ROW=256 PBK_KUBO_WN=4" "PBK_KUBO_ROW=256 PBK_KUBO_WN=2" "PBK_KUBO_ROW=256 PBK_KUBO_WN=4 PBK_KUBO_WAVES=4" "PBK_KUBO_ROW=256 PBK_KUBO_WN=4 PBK_KUBO_WAVES=16" "PBK_KUBO_ROW=256 PBK_KUBO_WN=4 PBK_KUBO_WAVES=32"; do
  echo "# $cfg" >> gpurun_out/r02_kubo_gemm_sweep.log
  env $cfg timeout 300 python tools/kubo_bench.py --vectors 4 --reps 1 >> gpurun_out/r02_kubo_gemm_sweep.log 2>&1
done
cat gpurun_out/r02_kubo_gemm_sweep.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kubo or conductivity" 2>&1 | tail -4
timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=128 PBK_RES=1,PBK_RES_ROW=32,PBK_RES_TILE=512 > gpurun_out/r02_sweep_cubic_res_v7.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos" gpurun_out/r02_sweep_cubic_res_v7.log | cut -c1-330
