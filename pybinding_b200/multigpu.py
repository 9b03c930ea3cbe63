"""Host-side plumbing for one-process-per-GPU runs (no reference counterpart: the reference is single-process,
thread-pool parallel over vector batches, cppcore/src/kpm/default/Compute.cpp:52-88).

Independent units (stochastic vectors of DOS / conductivity, LDOS sites) are sharded over ranks in contiguous
blocks (`pbk_shard`); the engine then needs exactly one collective, the ncclAllReduce of the moment accumulator,
for which every rank's `pbk_ctx` joins one NCCL communicator.  Rendezvous (who is rank 0, how 128 bytes travel)
is the launcher's business: any object with torch.distributed's `broadcast` works, on `nccl` or `gloo`.
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["shard", "broadcast_unique_id", "attach"]


def shard(total, world_size, rank):
    """`(first, count)`: the contiguous block of `total` units owned by `rank` (remainder to the lowest ranks)"""
    first, count = C.c_int32(), C.c_int32()
    status = _lib.load().pbk_shard(int(total), int(world_size), int(rank), C.byref(first), C.byref(count))
    if status != _lib.OK:
        raise ValueError("invalid shard request: total={}, world_size={}, rank={}".format(total, world_size, rank))
    return first.value, count.value


def broadcast_unique_id(dist, rank, device="cpu", make_id=None):
    """Create the 128-byte ncclUniqueId on rank 0 and broadcast it with `dist` (torch.distributed)"""
    import torch
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        if make_id is None:
            buf = C.create_string_buffer(128)
            status = _lib.load().pbk_comm_unique_id(buf)
            if status != _lib.OK:
                raise _lib.PbkError((_lib.load().pbk_last_error(None) or b"pbk_comm_unique_id failed").decode())
            raw = buf.raw
        else:
            raw = make_id()
        uid = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    dist.broadcast(uid, 0)
    return bytes(np.asarray(uid.cpu()).tobytes())


def attach(kpm, dist, rank, world_size, device):
    """Join `kpm`'s context to a communicator spanning all ranks of `dist`; afterwards the sharded entry points
    (calc_dos, calc_conductivity, calc_ldos, moments_*) split their units over the ranks and all-reduce the moments."""
    if world_size <= 1:
        return kpm
    uid = broadcast_unique_id(dist, rank, device)
    impl = getattr(kpm, "impl", kpm)
    impl.comm_init(world_size, rank, uid)
    return kpm
