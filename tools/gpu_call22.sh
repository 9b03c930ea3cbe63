#!/bin/bash
# GPU pass 22: where the host side of kpm.model = model + first use goes (PBK_TIMING)
mkdir -p gpurun_out
PBK_TIMING=1 timeout 600 python - > gpurun_out/host_timing.log 2>&1 <<'PY'
import time, numpy as np, bench
import pybinding_b200 as pb
w = bench.WORKLOADS["graphene_1000nm_c64_dos"]
model = bench.build_model(w)
k = pb.kpm(model, energy_range=w["energy_range"], silent=True)
for i in range(3):
    t0 = time.perf_counter(); k.model = model; t1 = time.perf_counter()
    m = k.impl.moments_dos(18, 8); t2 = time.perf_counter()
    print("iteration", i, "set model %.3f s, first moments call (build + 8 steps) %.3f s, hamiltonian_s %.3f" % (t1 - t0, t2 - t1, k.stats.hamiltonian_time), flush=True)
PY
cat gpurun_out/host_timing.log
