#!/bin/bash
# r02 GPU pass 14: resident-tile kernel with cp.async halo chunks (parity + sweep); DMMA main-loop probe; cuBLAS DGEMM for reference
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "resident" > gpurun_out/r02_pytest_res_v6.log 2>&1; tail -4 gpurun_out/r02_pytest_res_v6.log
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_RES=1 PBK_RES=1,PBK_RES_TILE=256 PBK_RES=1,PBK_RES_TILE=512 PBK_RES=1,PBK_RES_STAGES=3 \
  PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=256,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_BUFS=2,PBK_RES_TILE=384,PBK_RES_CTAS=2 \
  PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=192 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=256,PBK_RES_CTAS=2 PBK_RES=1,PBK_RES_ROW=128,PBK_RES_TILE=128,PBK_RES_BUFS=2,PBK_RES_CTAS=2 \
  > gpurun_out/r02_sweep_cubic_res_v8.log 2>&1
grep -v "cluster_order\|build_device\|set_hamiltonian\|calc_dos\|moments_dos" gpurun_out/r02_sweep_cubic_res_v8.log | cut -c1-250
tools/probe/build/gemm_loop_probe > gpurun_out/r02_gemm_loop_probe.jsonl 2>&1; cat gpurun_out/r02_gemm_loop_probe.jsonl
python - <<'PY' > gpurun_out/r02_cublas_dgemm.log 2>&1
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
for (m, k) in ((514, 1526122), (640, 1526122), (8192, 8192)):
    n = m if k > 8192 else 8192
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(n, k, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b.T
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): c = a @ b.T
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print({"cublas_dgemm": [m, n, k], "ms": round(ms, 3), "tflops": round(2.0 * m * n * k / ms / 1e9, 2)})
PY
cat gpurun_out/r02_cublas_dgemm.log
