"""Host side of the light-cone LDOS path (engine.cu: light_cone): the truncated breadth-first walk must visit the sites in
exactly the order of the reference's full relabelling (OptimizedHamiltonian::create_reordered, restated in the oracle
and pinned there to the known answers of cppcore/tests/test_kpm.cpp:33-171), shell by shell, and stop after `depth` shells."""
import ctypes

import numpy as np
import pytest

import pybinding_b200 as pb
from oracle.oracle import OracleKPM
from pybinding_b200 import _lib


def light_cone(h, src, depth):
    lib = _lib.load()
    n = h.shape[0]
    indptr = np.ascontiguousarray(h.indptr, np.int32)
    indices = np.ascontiguousarray(h.indices, np.int32)
    queue = np.empty(n, np.int32)
    borders = np.empty(depth + 1, np.int32)
    nq, nb, ex = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
    rc = lib.pbk_light_cone(ctypes.c_int64(n), _lib.ptr(indptr), _lib.ptr(indices), int(src), int(depth), _lib.ptr(queue),
                            ctypes.byref(nq), _lib.ptr(borders), ctypes.byref(nb), ctypes.byref(ex))
    assert rc == 0
    return queue[:nq.value].copy(), borders[:nb.value].copy(), bool(ex.value)


@pytest.mark.parametrize("onsite", [0.0, 0.3])
def test_walk_order_equals_the_reference_relabelling(onsite):
    model = pb.graphene_rectangle(6.0, dtype=np.float64, onsite=onsite)      # with and without a stored diagonal
    h = model.hamiltonian.tocsr()
    n = h.shape[0]
    ref = OracleKPM(h, energy_range=(-9, 9))
    for src in (0, n // 3, model.system.find_nearest([0.3, -0.2]), n - 1):
        full = ref.optimize_for([src], [src])
        rmap, slices = full["reorder_map"], full["slices"]
        for depth in (0, 1, 2, 7, 40, 10_000):
            queue, borders, exhausted = light_cone(h, src, depth)
            assert np.array_equal(rmap[queue], np.arange(queue.size))          # position i holds the site the reference puts at i
            assert np.array_equal(borders, slices[:borders.size])              # shell borders = the reference's slice map
            assert exhausted == (depth >= slices.size)
            assert queue.size == (n if exhausted else slices[depth])
            assert borders.size == min(depth + 1, slices.size)


def test_cone_is_small_on_a_large_sample():
    model = pb.graphene_rectangle(80.0, dtype=np.float32)                      # 245 k sites
    h = model.hamiltonian.tocsr()
    src = model.system.find_nearest([1.0, 2.0])                                # away from the edges
    queue, borders, exhausted = light_cone(h, src, 64)
    assert not exhausted and borders.size == 65 and borders[:4].tolist() == [1, 4, 10, 19]
    assert queue.size == borders[-1] == 1 + 3 * 64 * 65 // 2                    # the honeycomb ball: 2.5 % of the sample
    assert len(set(queue.tolist())) == queue.size
