#!/usr/bin/env python
"""Diagnostic 2: which knob removes the run-to-run differences of the few-lane staged kernel at 38 M sites?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pybinding_b200 as pb

KN = ("PBK_MT_SEQUENTIAL", "PBK_BULK", "PBK_DEVBUILD", "PBK_TILE", "PBK_XS", "PBK_BPSM", "PBK_GRAPH")
def run(model, er, M, R, reps, **env):
    for k in KN: os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    kpm = pb.kpm(model, energy_range=er, silent=True)
    return [kpm.impl.moments_dos(M, R) for _ in range(reps)]

er = (-8.5, 8.5)
for dtype in (np.complex64, np.float32):
    model = pb.graphene_rectangle(1000.0, magnetic_field=10.0 if dtype == np.complex64 else 0.0, dtype=dtype)
    M = 18
    lanes = (2, 4, 8, 16) if dtype == np.complex64 else (4, 8, 16)
    for R in lanes:
        ref = run(model, er, M, R, 1, PBK_BULK=0)[0]
        scale = np.abs(ref).max()
        for env in ({}, {"PBK_XS": 0}, {"PBK_BULK": 2}, {"PBK_BULK": 8}, {"PBK_BPSM": 1}, {"PBK_BPSM": 2}):
            outs = run(model, er, M, R, 4, **env)
            errs = [float(np.abs(o - ref).max() / scale) for o in outs]
            bad = sorted({int(i) for o in outs for i in np.flatnonzero(np.abs(o - ref) / scale > 1e-9)})
            print(json.dumps(dict(dtype=np.dtype(dtype).name, R=R, env=env, errs=errs, bad_moments=bad)), flush=True)
