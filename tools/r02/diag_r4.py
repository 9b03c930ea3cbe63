#!/usr/bin/env python
"""Diagnostic: DOS moments with few lanes on large graphene systems, GPU variants against the hp oracle."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pybinding_b200 as pb
from oracle.oracle import OracleKPM, hardware_threads

def gpu(model, er, M, R, **env):
    for k in ("PBK_MT_SEQUENTIAL", "PBK_BULK", "PBK_DEVBUILD", "PBK_TILE"): os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    kpm = pb.kpm(model, energy_range=er, silent=True)
    a = kpm.impl.moments_dos(M, R)
    b = kpm.impl.moments_dos(M, R)
    return a, b, kpm

for size in (float(a) for a in sys.argv[1:] or ["300", "1000"]):
    model = pb.graphene_rectangle(size, magnetic_field=10.0, dtype=np.complex64)
    n = model.hamiltonian.shape[0]
    er = (-8.5, 8.5)
    ref = OracleKPM(model.hamiltonian, energy_range=er, hp=True, num_threads=hardware_threads())
    M = 10
    for R in (1, 4, 8):
        t0 = time.time(); hp = ref.dos_moments(M, R); t_or = time.time() - t0
        scale = np.abs(hp).max()
        for env in ({}, {"PBK_MT_SEQUENTIAL": 1}, {"PBK_BULK": 0}, {"PBK_TILE": -1}):
            a, b, kpm = gpu(model, er, M, R, **env)
            print(json.dumps(dict(size=size, n=n, R=R, env=env, rel=float(np.abs(a - hp).max() / scale), repeat_equal=bool(np.array_equal(a, b)),
                                  absdiff=[float(x) for x in np.abs(a - hp)], oracle_s=round(t_or, 1))), flush=True)
    # starters: first vector, GPU stream against the oracle's
    kpm = pb.kpm(model, energy_range=er, silent=True)
    g = kpm.impl.random_vectors(2)
    o = ref.random_vectors(2)
    d = np.abs(g - o)
    print(json.dumps(dict(size=size, starters_max_abs=float(d.max()), mismatches_gt_1e5=int((d > 1e-5).sum()),
                          first_bad=[int(i) for i in np.argwhere(d > 1e-5)[:5].ravel()])), flush=True)
