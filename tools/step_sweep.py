#!/usr/bin/env python
"""Sweep layout / launch parameters of the fused Chebyshev step kernel (env overrides read by the engine).

    python tools/step_sweep.py [--workload graphene_200nm_c64_dos] [--moments 130] [--vectors 64] CONFIG...
CONFIG = comma-separated KEY=VALUE pairs, e.g.  PBK_TILE=-1  PBK_TILE=256,PBK_TPB=512
Prints one line per config: ms per Chebyshev step (of one pass of `batch` vectors), algorithmic GB/s, fraction of the measured HBM peak.
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
import pybinding_b200 as pb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="graphene_200nm_c64_dos")
    ap.add_argument("--moments", type=int, default=130)
    ap.add_argument("--vectors", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("configs", nargs="*", default=["PBK_TILE=-1"])
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    R = args.vectors or w["vectors"]
    t0 = time.time()
    model = bench.build_model(w)
    print("# model {} built in {:.1f} s: n={} nnz={}".format(args.workload, time.time() - t0, model.hamiltonian.shape[0],
                                                          model.hamiltonian.nnz), flush=True)
    peak = bench.measured_peak()[0]
    ref = None
    for cfg in args.configs:
        env = dict(kv.split("=") for kv in cfg.split(",") if kv)
        max_batch = int(env.pop("MB", 0))   # pseudo-key: vectors per pass
        for k in [k for k in os.environ if k.startswith("PBK_") and k != "PBK_TIMING"]:   # every config starts from the defaults
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = time.time()
        kpm = pb.kpm(model, energy_range=w["energy_range"], silent=True, max_batch=max_batch)
        mom = None
        best = None
        sampler = bench.ClockSampler(0)
        sampler.start()
        for _ in range(args.reps + 1):
            mom = kpm.impl.moments_dos(args.moments, R)
            s = kpm.stats
            steps = s.step_launches
            ms = s.step_ms / steps
            best = ms if best is None else min(best, ms)
        clocks = sampler.stop()
        gbs = s.step_bytes / steps / (best * 1e-3) / 1e9
        if ref is None:
            ref = mom
        err = float(np.abs(mom - ref).max() / np.abs(ref).max())
        print(json.dumps(dict(config=cfg, ms_per_step=round(best, 4), vectors_per_ms=round(s.batch / best, 2), algorithmic_gbs=round(gbs, 1),
                              frac=round(gbs / peak, 4), hamiltonian_s=round(s.hamiltonian_time, 2),
                              starter_ms=round(s.starter_ms, 1), batch=s.batch, rel_diff_vs_first=err, sm_mhz=clocks.get("sm_mhz"),
                              power_w=clocks.get("power_w"), res_launches=int(s.res_launches), bulk_launches=int(s.bulk_launches),
                              total_s=round(time.time() - t0, 1))), flush=True)
        del kpm


if __name__ == "__main__":
    main()
