#!/usr/bin/env python
"""Small cases that touch every kernel added late in round 1 (two-step kernel, graph replay, light-cone LDOS, Green's):
run under compute-sanitizer (memcheck / racecheck / synccheck) by tools/gpu_call15.sh."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pybinding_b200 as pb
def run(env, fn):
    for k in ("PBK_PAIR", "PBK_GRAPH", "PBK_CONE", "PBK_PAIR_MINB"): os.environ.pop(k, None)
    os.environ.update(env)
    return fn()
m32 = pb.graphene_rectangle(12.0, dtype=np.complex64, magnetic_field=300.0)
m64 = pb.graphene_rectangle(10.0, dtype=np.float64, onsite=0.2)
cub = pb.cubic_anderson(12, disorder=2.0, dtype=np.float32)
def dos(model, er, M, R):
    return pb.kpm(model, energy_range=er, silent=True).impl.moments_dos(M, R)
a = run({"PBK_PAIR": "0", "PBK_GRAPH": "0"}, lambda: dos(m32, (-9, 9), 34, 8))
b = run({"PBK_PAIR": "1", "PBK_GRAPH": "0"}, lambda: dos(m32, (-9, 9), 34, 8))
c = run({"PBK_PAIR": "2", "PBK_PAIR_MINB": "2"}, lambda: dos(cub, (-8.2, 8.2), 34, 8))
d = run({"PBK_PAIR": "1"}, lambda: dos(m64, (-9, 9), 34, 6))
g = run({}, lambda: [dos(m64, (-9, 9), 34, 1) for _ in range(2)])
k = pb.kpm(m64, energy_range=(-9, 9), silent=True)
fn = m64.system.find_nearest
l1 = k.impl.moments_ldos(66, [fn([0, 0])])
l2 = k.impl.moments_ldos(18, [fn([0, 0]), fn([3, 3]), fn([-4, 2])])
gr = k.impl.moments_greens(34, fn([0, 0]), [fn([1, 1]), fn([2, -2])])
print("pair diff", float(abs(a - b).max() / abs(a).max()), "ok")
