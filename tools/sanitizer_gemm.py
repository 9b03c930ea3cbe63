#!/usr/bin/env python
"""One Kubo-Bastin moment matrix of a small system, for compute-sanitizer runs of the GEMM alone:
    compute-sanitizer --tool racecheck python tools/sanitizer_gemm.py M [dtype]
M = 128: one full tile, every warp active; 134: remainder folded into the tile; 18: a single edge tile (one active warp);
200: 2 x 2 tiles with edge tiles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pybinding_b200 as pb

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dtype = np.dtype(sys.argv[2]) if len(sys.argv) > 2 else np.dtype(np.float64)
model = pb.graphene_rectangle(6.0, dtype=dtype, onsite=0.2, magnetic_field=300.0 if dtype.kind == "c" else 0.0)
k = pb.kpm(model, energy_range=(-9, 9), silent=True)
mu = k.impl.moments_kubo(M, model.system.x, model.system.y, 1)
print("M", M, dtype.name, "sites", model.hamiltonian.shape[0], "trace", float(np.trace(mu).real), "gemm launches ok")
