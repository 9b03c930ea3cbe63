#!/usr/bin/env python
"""Checks bench.py's CPU baseline (a bounded sample, extrapolated linearly from the per-step cost and the fixed cost per
wave) against a FULL-LENGTH run of the same CPU port on a system small enough to finish: graphene 300 x 300 nm
(3.4 M sites), complex64, 64 vectors.  No GPU needed.

    python tools/validate_cpu_baseline.py [--moments 514] [--size 300]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.oracle import OracleKPM, hardware_threads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--moments", type=int, default=514)
    ap.add_argument("--size", type=float, default=300.0)
    ap.add_argument("--vectors", type=int, default=64)
    args = ap.parse_args()
    w = dict(kind="graphene", size=args.size, field=10.0, dtype="complex64", moments=args.moments, vectors=args.vectors,
             energy_range=(-8.5, 8.5), text="validation")
    model = bench.build_model(w)
    nnz = model.hamiltonian.nnz
    threads = hardware_threads()
    sample, seconds = bench.cpu_sample(model, w, threads)
    ref = OracleKPM(model.hamiltonian, energy_range=w["energy_range"], num_threads=threads, hp=False)
    t0 = time.perf_counter()
    ref.dos_moments(args.moments, args.vectors)          # the whole job, real MT19937 starters
    full = time.perf_counter() - t0
    full_value = nnz * args.moments * args.vectors / full
    print(json.dumps(dict(sites=int(model.hamiltonian.shape[0]), nnz=int(nnz), moments=args.moments, vectors=args.vectors, threads=threads,
                          sample_seconds=round(seconds, 2), estimated_job_seconds=round(sample["extrapolated_job_seconds"], 2),
                          estimated_value=sample["value"], full_run_seconds=round(full, 2), full_run_value=full_value,
                          estimate_over_full=round(sample["value"] / full_value, 4))))


if __name__ == "__main__":
    main()
