#!/bin/bash
# GPU pass (round 1, v2): staged (bulk-copy) step kernel -- quick check, parity tests, sweeps, Kubo GEMM numbers + ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia-smi.txt 2>&1
# 1. quick check of the staged kernel (guards the rest of the call against a hang)
timeout 180 python - > gpurun_out/quick.log 2>&1 <<'PY'
import os, numpy as np
import pybinding_b200 as pb
from oracle.oracle import OracleKPM
model = pb.graphene_rectangle(20.0, dtype=np.complex64, magnetic_field=200.0)
ref = OracleKPM(model.hamiltonian, energy_range=(-9, 9), hp=True).dos_moments(66, 8)
for env in ({"PBK_BULK": "0"}, {}, {"PBK_XS": "0"}):
    for k in ("PBK_BULK", "PBK_XS"): os.environ.pop(k, None)
    os.environ.update(env)
    kpm = pb.kpm(model, energy_range=(-9, 9), silent=True)
    m = kpm.impl.moments_dos(66, 8)
    s = kpm.stats
    print(env, "err", float(np.abs(m - ref).max() / np.abs(ref).max()), "bulk", s.bulk_launches, "steps", s.step_launches, flush=True)
PY
echo "quick exit $?" >> gpurun_out/quick.log
if ! grep -q "quick exit 0" gpurun_out/quick.log; then export PBK_BULK=0; echo "STAGED KERNEL FAILED: falling back to PBK_BULK=0" >> gpurun_out/quick.log; fi
# 2. parity tests
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
# 3. sweeps at benchmark size
timeout 1200 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 \
  PBK_BULK=0 PBK_BULK=4 PBK_BULK=3 PBK_BULK=6 PBK_BULK=2 PBK_BULK=4,PBK_XS=0 PBK_BULK=8,PBK_XS=0 PBK_BULK=4,PBK_BPSM=3 PBK_BULK=6,PBK_BPSM=3 \
  PBK_BULK=4,PBK_BPSM=5 PBK_BULK=4,PBK_TILE=512 PBK_BULK=4,PBK_TILE=128 PBK_BULK=8,PBK_XS=0,PBK_TILE=512 > gpurun_out/sweep_full.log 2>&1
timeout 900 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 \
  PBK_BULK=0 PBK_BULK=4 PBK_BULK=4,PBK_TILE=1024 PBK_BULK=4,PBK_TILE=4096 PBK_BULK=0,PBK_TILE=4096 PBK_BULK=8,PBK_XS=0,PBK_TILE=4096 PBK_BULK=4,PBK_TILE=16384 > gpurun_out/sweep_cubic.log 2>&1
# 4. Kubo-Bastin: stacks + GEMM at configs[3] size, then one ncu capture of the GEMM
timeout 600 python tools/kubo_bench.py --reps 1 > gpurun_out/kubo_200nm.json 2> gpurun_out/kubo_200nm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kubo_gemm -c 1 -f -o gpurun_out/kubo_gemm \
  python tools/kubo_bench.py --size 100 --reps 0 > gpurun_out/ncu_kubo.log 2>&1
cat gpurun_out/quick.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_full.log gpurun_out/sweep_cubic.log gpurun_out/kubo_200nm.json; tail -3 gpurun_out/kubo_200nm.err gpurun_out/ncu_kubo.log
