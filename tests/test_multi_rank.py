"""World-size-2 `gloo` tests of the multi-rank host logic (CPU only; the NCCL path itself runs in bench.py --gpus N).

What is checked: the unit partition every sharded entry point uses (`pbk_shard`), the unique-id rendezvous, and the
reduction semantics -- per-rank sums of per-vector moments over contiguous shards of the reference's single random
stream, one all-reduce, one division by num_random -- against the oracle's single-process result
(reference: BatchAccumulator, cppcore/src/kpm/Moments.cpp:7-49).
"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch
    import torch.distributed as dist
    import pybinding_b200 as pb
    from pybinding_b200 import multigpu
    from oracle.oracle import OracleKPM

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)

    # 1. partition: contiguous, disjoint, covering, remainder to the lowest ranks
    out = dict(rank=rank)
    for total in (0, 1, 2, 7, 64):
        first, count = multigpu.shard(total, world, rank)
        t = torch.tensor([first, count], dtype=torch.int64)
        gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, t)
        spans = [tuple(g.tolist()) for g in gathered]
        pos = 0
        for f, c in spans:
            assert f == pos and c >= 0, (total, spans)
            pos += c
        assert pos == total, (total, spans)
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
        assert sorted((c for _, c in spans), reverse=True) == [c for _, c in spans]

    # 2. rendezvous of the 128-byte communicator id (payload made up here: no NCCL on a CPU box)
    payload = bytes(range(128))
    uid = multigpu.broadcast_unique_id(dist, rank, "cpu", make_id=lambda: payload)
    assert uid == payload

    # 3. sharded stochastic trace == single-process trace
    model = pb.graphene_rectangle(6, onsite=0.25, magnetic_field={field}, dtype=np.dtype({dtype!r}))
    ref = OracleKPM(model.hamiltonian, energy_range=(-9, 9), kernel="dirichlet", hp=True)
    M, R = 34, 5
    vectors = ref.random_vectors(R)                      # the reference's stream: vector j = draws [j*N*w, (j+1)*N*w)
    first, count = multigpu.shard(R, world, rank)
    local = np.zeros(M, np.complex128)
    for j in range(first, first + count):
        local += ref.moments(M, vectors[j])              # <r_j| T_n |r_j>, mu_0 halved, dirichlet: undamped
    t = torch.from_numpy(np.ascontiguousarray(local.view(np.float64)))
    dist.all_reduce(t)                                   # THE one collective of the path
    mean = t.numpy().view(np.complex128) / R
    expected = ref.dos_moments(M, R)
    err = float(np.abs(mean - expected).max() / np.abs(expected).max())
    assert err < 1e-12, err
    out["err"] = err
    dist.barrier()
    dist.destroy_process_group()
    print("RESULT " + json.dumps(out))
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dtype,field", [("float64", 0.0), ("complex128", 50.0)])
def test_world_size_2_gloo(tmp_path, dtype, field):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, dtype=dtype, field=field))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outputs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outputs.append(out)
    for p, out in zip(procs, outputs):
        assert p.returncode == 0, out[-3000:]
        assert "RESULT " in out, out[-3000:]


def test_shard_rejects_bad_arguments():
    from pybinding_b200 import multigpu
    with pytest.raises(ValueError):
        multigpu.shard(4, 2, 2)
    with pytest.raises(ValueError):
        multigpu.shard(4, 0, 0)
    assert multigpu.shard(5, 2, 0) == (0, 3) and multigpu.shard(5, 2, 1) == (3, 2)
