#!/usr/bin/env python
"""LDOS sites sharded over the ranks of a torchrun launch: every rank returns the full table (one ncclAllReduce),
compared on rank 0 with an un-sharded run of the same sites."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import pybinding_b200 as pb
from pybinding_b200 import multigpu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = pb.graphene_rectangle(200.0, dtype=np.complex128, magnetic_field=10.0, disorder=0.5, disorder_seed=0)
fn = model.system.find_nearest
xs = np.linspace(-80, 80, 6)
sites = [fn([x, y]) for x in xs for y in xs]
M = 514
kpm = pb.kpm(model, energy_range=(-8.8, 8.8), silent=True, device=local)
multigpu.attach(kpm, dist, rank, world, "cuda")
kpm.impl.moments_ldos(M, sites[:world])   # first collective of the communicator (NCCL sets up its channels here), untimed
dist.barrier()
t0 = time.perf_counter()
table = kpm.impl.moments_ldos(M, sites)
dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    single = pb.kpm(model, energy_range=(-8.8, 8.8), silent=True, device=local)
    single.impl.moments_ldos(M, sites[:1])
    t0 = time.perf_counter()
    ref = single.impl.moments_ldos(M, sites)
    dt1 = time.perf_counter() - t0
    err = float(np.abs(table - ref).max() / np.abs(ref).max())
    print("ldos {} sites, M={}, {} ranks: {:.3f} s (1 rank: {:.3f} s), max rel diff {:.2e}, groups per rank {}".format(
        len(sites), M, world, dt, dt1, err, kpm.stats.num_batches))
    assert err < 1e-12
dist.destroy_process_group()
