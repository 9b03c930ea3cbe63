#!/usr/bin/env python
"""Host -> device upload of a CSR-sized array (1.5 GB) from PAGEABLE memory (cudaMemcpyAsync stages it inside the driver)
against the engine's route (parallel memcpy into a page-locked mirror, then DMA).  Decides whether `set_hamiltonian` may
skip the mirror.  Run on the GPU box."""
import time
import numpy as np
import torch

n = 1_526_000_000 // 4
src = np.arange(n, dtype=np.int32)
src[::4096] += 1                      # touched
dst = torch.empty(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for rep in range(4):
    t0 = time.perf_counter()
    dst.copy_(torch.from_numpy(src))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("pageable numpy -> device: %.3f s  %.1f GB/s" % (dt, src.nbytes / dt / 1e9))
pinned = torch.empty(n, dtype=torch.int32).pin_memory()
for rep in range(3):
    t0 = time.perf_counter()
    pinned.numpy()[:] = src           # single-threaded mirror
    t1 = time.perf_counter()
    dst.copy_(pinned, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("mirror (1 thread) %.3f s + pinned DMA %.3f s  (%.1f GB/s)" % (t1 - t0, t2 - t1, src.nbytes / (t2 - t1) / 1e9))
