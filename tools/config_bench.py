#!/usr/bin/env python
"""End-to-end timings of BASELINE.json configs[2] and configs[3] through the public API (host buffers in, curves out).

    python tools/config_bench.py ldos   [--size 500] [--sites 256] [--broadening 0.02]
    python tools/config_bench.py greens [--size 500] [--sites 256] [--broadening 0.02]
    python tools/config_bench.py sigma  [--size 200] [--vectors 4] [--points 1000]
Prints one JSON line per run (wall seconds of the API call, engine stats).
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


def stats_dict(s):
    return dict(num_moments=int(s.num_moments), batch=int(s.batch), num_batches=int(s.num_batches), nnz=int(s.nnz),
                opt_nnz=int(s.opt_nnz), step_launches=int(s.step_launches), bulk_launches=int(s.bulk_launches),
                step_ms=round(s.step_ms, 2), gemm_ms=round(s.gemm_ms, 2),
                gemm_tflops=round(s.gemm_flops / (s.gemm_ms * 1e-3) / 1e12, 2) if s.gemm_ms else None,
                moments_device_ms=round(s.moments_device_ms, 2), hamiltonian_s=round(s.hamiltonian_time, 2),
                step_gbs=round(s.step_bytes / (s.step_ms * 1e-3) / 1e9, 1) if s.step_ms else None, eps=float(s.eps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["ldos", "greens", "sigma"])
    ap.add_argument("--size", type=float, default=0)
    ap.add_argument("--sites", type=int, default=256)
    ap.add_argument("--broadening", type=float, default=0.02)
    ap.add_argument("--vectors", type=int, default=4)
    ap.add_argument("--points", type=int, default=1000)
    args = ap.parse_args()
    t0 = time.time()
    if args.what in ("ldos", "greens"):
        size = args.size or 500.0
        model = pb.graphene_rectangle(size, dtype=np.complex128, magnetic_field=10.0, disorder=0.5, disorder_seed=0)
        kpm = pb.kpm(model, energy_range=(-8.8, 8.8), silent=True)
        g = int(round(np.sqrt(args.sites)))
        xs = np.linspace(-0.4 * size, 0.4 * size, g)
        sites = [model.system.find_nearest([x, y]) for x in xs for y in xs][:args.sites]
        energy = np.linspace(-1, 1, 500)
        setup = time.time() - t0
        t1 = time.time()
        if args.what == "ldos":
            out = kpm.impl._ldos_indices(sites, energy, args.broadening)   # core.ldos(indices): E x sites
            check = float(out.sum())
            text = "graphene {:g}x{:g} nm + onsite disorder + Peierls field, complex128, LDOS at {} sites".format(size, size, len(sites))
        else:
            out = kpm.calc_greens(sites[len(sites) // 2], sites, energy, args.broadening)
            check = float(np.abs(np.asarray(out)).sum())
            text = "graphene {:g}x{:g} nm + onsite disorder + Peierls field, complex128, Green's i -> {} sites".format(size, size, len(sites))
    else:
        size = args.size or 200.0
        model = pb.graphene_rectangle(size, dtype=np.float64)
        kpm = pb.kpm(model, energy_range=(-9, 9), kernel=pb.lorentz_kernel(), silent=True)
        a, _ = kpm.scaling_factors
        broadening = a * 4.0 / 512.5     # Lorentz lambda = 4 -> 514 moments
        mu = np.linspace(-1, 1, 101)
        setup = time.time() - t0
        t1 = time.time()
        out = kpm.calc_conductivity(mu, broadening, 300.0, "xx", num_random=args.vectors, num_points=args.points)
        t_xx = time.time() - t1
        sxx = stats_dict(kpm.stats)
        t1 = time.time()
        out2 = kpm.calc_conductivity(mu, broadening, 300.0, "xy", num_random=args.vectors, num_points=args.points)
        print(json.dumps(dict(workload="graphene {:g}x{:g} nm float64 calc_conductivity xx, {} vectors, {} points".format(size, size, args.vectors, args.points),
                              sites=int(model.hamiltonian.shape[0]), wall_s=round(t_xx, 3), checksum=float(np.abs(out.data).sum()), stats=sxx)))
        out, check = out2, float(np.abs(out2.data).sum())
        text = "graphene {:g}x{:g} nm float64 calc_conductivity xy, {} vectors, {} points".format(size, size, args.vectors, args.points)
    wall = time.time() - t1
    print(json.dumps(dict(workload=text, sites=int(model.hamiltonian.shape[0]), setup_s=round(setup, 1), wall_s=round(wall, 3),
                          checksum=check, stats=stats_dict(kpm.stats))))


if __name__ == "__main__":
    main()
