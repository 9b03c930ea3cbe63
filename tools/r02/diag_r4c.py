#!/usr/bin/env python
"""Diagnostic 3: does handing the stage back later (PBK_RELEASE=1) or a proxy fence (2) remove the few-lane differences?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pybinding_b200 as pb

KN = ("PBK_BULK", "PBK_XS", "PBK_BPSM", "PBK_RELEASE", "PBK_RES")
def run(model, er, M, R, reps, **env):
    for k in KN: os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    kpm = pb.kpm(model, energy_range=er, silent=True)
    return [kpm.impl.moments_dos(M, R) for _ in range(reps)], kpm.stats.step_ms / max(kpm.stats.step_launches, 1)

er = (-8.5, 8.5)
for dtype, lanes in ((np.complex64, (2, 4, 64)), (np.float32, (4,))):
    model = pb.graphene_rectangle(1000.0, magnetic_field=10.0 if dtype == np.complex64 else 0.0, dtype=dtype)
    for R in lanes:
        M = 18 if R < 64 else 10
        ref = run(model, er, M, R, 1, PBK_BULK=0)[0][0]
        scale = np.abs(ref).max()
        for env in ({}, {"PBK_RELEASE": 1}, {"PBK_RELEASE": 2}, {"PBK_XS": 0}, {"PBK_BULK": 3}, {"PBK_BULK": 3, "PBK_RELEASE": 1}):
            outs, ms = run(model, er, M, R, 6 if R < 64 else 2, **env)
            errs = [float(np.abs(o - ref).max() / scale) for o in outs]
            print(json.dumps(dict(dtype=np.dtype(dtype).name, R=R, env=env, ms_per_step=round(ms, 4), max_err=max(errs), errs=errs)), flush=True)
