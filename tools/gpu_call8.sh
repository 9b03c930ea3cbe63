#!/bin/bash
# GPU pass 8: first run of the two-step kernel (cheb_pair_bulk): quick check, its parity tests, sweep at benchmark size.
mkdir -p gpurun_out
timeout 240 python - > gpurun_out/pair_quick.log 2>&1 <<'PY'
import os, numpy as np
import pybinding_b200 as pb
from oracle.oracle import OracleKPM
model = pb.graphene_rectangle(20.0, dtype=np.complex64, magnetic_field=200.0)
ref = OracleKPM(model.hamiltonian, energy_range=(-9, 9), hp=True).dos_moments(66, 8)
for env in ({"PBK_PAIR": "0"}, {"PBK_PAIR": "1"}, {"PBK_PAIR": "1", "PBK_PAIR_MINB": "2"}):
    for k in ("PBK_PAIR", "PBK_PAIR_MINB"): os.environ.pop(k, None)
    os.environ.update(env)
    kpm = pb.kpm(model, energy_range=(-9, 9), silent=True)
    m = kpm.impl.moments_dos(66, 8)
    s = kpm.stats
    print(env, "err", float(np.abs(m - ref).max() / np.abs(ref).max()), "pair", s.pair_launches, "bulk", s.bulk_launches, "steps", s.step_launches, flush=True)
PY
echo "quick exit $?" >> gpurun_out/pair_quick.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "pair" > gpurun_out/pytest_pair.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pair.log
timeout 900 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 \
  PBK_PAIR=0 PBK_PAIR=1 PBK_PAIR=1,PBK_PAIR_STAGES=6 PBK_PAIR=1,PBK_PAIR_STAGES=3 PBK_PAIR=0,MB=32 PBK_PAIR=1,MB=32 PBK_PAIR=1,MB=32,PBK_PAIR_MINB=2 \
  PBK_PAIR=1,MB=32,PBK_PAIR_STAGES=3 PBK_PAIR=1,MB=32,PBK_PAIR_STAGES=6,PBK_PAIR_MINB=2 PBK_PAIR=1,MB=16 PBK_PAIR=1,MB=32,PBK_TILE=512 PBK_PAIR=1,MB=16,PBK_TILE=512 \
  > gpurun_out/sweep_pair_full.log 2>&1
cat gpurun_out/pair_quick.log; tail -n 15 gpurun_out/pytest_pair.log; cat gpurun_out/sweep_pair_full.log
