#!/usr/bin/env python
"""Kubo-Bastin moment matrix (K3 stacks + K4 GEMM) on BASELINE configs[3]: graphene 200x200 nm, float64, M = 514.

    python tools/kubo_bench.py [--size 200] [--moments 514] [--vectors 1] [--dtype float64] [--direction xx]
Prints one JSON line: recursion (stack) time, GEMM time, GEMM TFLOP/s (2*M^2*N real flops per vector; x4 complex).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pybinding_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=float, default=200.0)
    ap.add_argument("--moments", type=int, default=514)
    ap.add_argument("--vectors", type=int, default=1)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--direction", default="xx")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    dtype = np.dtype(args.dtype)
    t0 = time.time()
    model = pb.graphene_rectangle(args.size, dtype=dtype, magnetic_field=10.0 if dtype.kind == "c" else 0.0)
    n = model.hamiltonian.shape[0]
    kpm = pb.kpm(model, energy_range=(-9, 9), kernel=pb.lorentz_kernel(), silent=True)
    left = model.system.x
    right = model.system.x if args.direction[1] == "x" else model.system.y
    best = None
    for _ in range(args.reps + 1):
        t1 = time.time()
        mu = kpm.impl.moments_kubo(args.moments, left, right, args.vectors)
        wall = time.time() - t1
        s = kpm.stats
        rec = dict(gemm_ms=s.gemm_ms / args.vectors, gemm_tflops=s.gemm_flops / (s.gemm_ms * 1e-3) / 1e12,
                   stack_ms=s.step_ms / args.vectors, moments_device_ms=s.moments_device_ms, wall_s=wall)
        if best is None or rec["gemm_ms"] < best["gemm_ms"]:
            best = rec
    best.update(workload="graphene {:g}x{:g} nm {} Kubo-Bastin {} M={} R={}".format(args.size, args.size, dtype.name,
                                                                                    args.direction, args.moments, args.vectors),
                n=int(n), mu_trace=float(np.trace(mu).real), setup_s=round(time.time() - t0, 1))
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in best.items()}))


if __name__ == "__main__":
    main()
