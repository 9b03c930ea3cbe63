#!/usr/bin/env python
"""Small cases that touch every kernel of the library (general / staged step kernels for fixed and generic ELL widths,
lane-aligned batches, graph replay, light-cone LDOS, Green's, Kubo-Bastin): run under compute-sanitizer
(memcheck / racecheck / synccheck)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pybinding_b200 as pb
def run(env, fn):
    for k in [k for k in os.environ if k.startswith("PBK_")]: os.environ.pop(k, None)
    os.environ.update(env)
    return fn()
m32 = pb.graphene_rectangle(12.0, dtype=np.complex64, magnetic_field=300.0)
m64 = pb.graphene_rectangle(10.0, dtype=np.float64, onsite=0.2)
cub = pb.cubic_anderson(12, disorder=2.0, dtype=np.float32)
def dos(model, er, M, R):
    return pb.kpm(model, energy_range=er, silent=True).impl.moments_dos(M, R)
from pybinding_b200 import synthetic as syn
nnn = syn.graphene_monolayer(pb.Rectangle(9.0), nearest_neighbors=2, dtype=np.float32)   # ELL width 10: generic-width kernels
a = run({"PBK_BULK": "0", "PBK_GRAPH": "0"}, lambda: dos(m32, (-9, 9), 34, 8))
b = run({"PBK_BULK": "4", "PBK_GRAPH": "0"}, lambda: dos(m32, (-9, 9), 34, 8))
c = run({}, lambda: dos(cub, (-8.2, 8.2), 34, 8))
d = run({}, lambda: dos(nnn, (-9.2, 9.6), 34, 12))
# batch caps that are not multiples of the 16-byte lane width (f64: 2 lanes, f32: 4 lanes)
e = run({}, lambda: pb.kpm(m64, energy_range=(-9, 9), silent=True, max_batch=3).impl.moments_dos(34, 7))
f = run({}, lambda: pb.kpm(cub, energy_range=(-8.2, 8.2), silent=True, max_batch=5).impl.moments_dos(34, 11))
g = run({}, lambda: [dos(m64, (-9, 9), 34, 1) for _ in range(2)])
k = pb.kpm(m64, energy_range=(-9, 9), silent=True)
fn = m64.system.find_nearest
l1 = k.impl.moments_ldos(66, [fn([0, 0])])
l2 = k.impl.moments_ldos(18, [fn([0, 0]), fn([3, 3]), fn([-4, 2])])
gr = k.impl.moments_greens(34, fn([0, 0]), [fn([1, 1]), fn([2, -2])])
ku = k.impl.moments_kubo(18, m64.system.x, m64.system.y, 2)
ku2 = k.impl.moments_kubo(134, m64.system.x, m64.system.x, 1)   # 128 + 6: remainder rows folded into the last tile of the GEMM
# resident-tile kernel (forced: the small lattice would not choose it), 16 float lanes = one 64-byte row
r1 = run({"PBK_RES": "2", "PBK_RES_TILE": "128"}, lambda: dos(cub, (-8.2, 8.2), 34, 16))
r0 = run({"PBK_RES": "0"}, lambda: dos(cub, (-8.2, 8.2), 34, 16))
print("resident vs staged diff", float(abs(r1 - r0).max() / abs(r0).max()))
print("staged vs general diff", float(abs(a - b).max() / abs(a).max()), "ok")
