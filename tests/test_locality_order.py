"""Host-side locality relabelling of the full-system layout (engine.cu: cluster_order, one- and two-level).

KPM results do not depend on the row order, so what is checked here is what the step kernel relies on: the order is a
permutation, tiles are compact (their one-ring halo is a fraction of the tile) and the two-level variant keeps the tiles
of a macro-block consecutive."""
import ctypes

import numpy as np
import pytest

import pybinding_b200 as pb
from pybinding_b200 import _lib


def order_of(h, tile, macro=None):
    lib = _lib.load()
    n = h.shape[0]
    indptr = np.ascontiguousarray(h.indptr, np.int32)
    indices = np.ascontiguousarray(h.indices, np.int32)
    order = np.empty(n, np.int32)
    if macro is None:
        rc = lib.pbk_locality_order(ctypes.c_int64(n), _lib.ptr(indptr), _lib.ptr(indices), tile, _lib.ptr(order))
    else:
        rc = lib.pbk_locality_order2(ctypes.c_int64(n), _lib.ptr(indptr), _lib.ptr(indices), tile, macro, _lib.ptr(order))
    assert rc == 0
    return order


def halo_fraction(h, order, tile):
    n = h.shape[0]
    rmap = np.empty(n, np.int64)
    rmap[order] = np.arange(n)
    rows = np.repeat(np.arange(n), np.diff(h.indptr))
    tr, nc = rmap[rows] // tile, rmap[h.indices]
    out = tr != nc // tile
    return len(np.unique(tr[out] * n + nc[out])) / n


@pytest.mark.parametrize("macro", [None, 0, 4, 16])
def test_graphene_order_is_a_compact_permutation(macro):
    h = pb.graphene_rectangle(40.0, dtype=np.float32).hamiltonian.tocsr()     # 61 k sites
    order = order_of(h, 256, macro)
    assert np.array_equal(np.sort(order), np.arange(h.shape[0]))
    assert halo_fraction(h, order, 256) < 0.4                     # natural (sublattice-major) order: 3.0
    if macro is None:
        assert np.array_equal(order, order_of(h, 256, 0))          # macro_tiles = 0 is the one-level order
        assert np.array_equal(order, order_of(h, 256))              # deterministic


def test_two_level_order_is_deterministic_and_blocks_are_closed_under_the_first_level(monkeypatch):
    h = pb.graphene_rectangle(60.0, dtype=np.float32).hamiltonian.tocsr()
    n, tile, macro = h.shape[0], 128, 8
    monkeypatch.setenv("PBK_COARSE", "1")                           # macro-blocks grown on the real graph
    a, b = order_of(h, tile, macro), order_of(h, tile, macro)
    assert np.array_equal(a, b)                                     # thread schedule does not leak into the result
    level1 = order_of(h, tile * macro)                              # the macro-blocks are the clusters of this order
    block = tile * macro
    for m in range(0, n, block):
        assert set(a[m:m + block]) == set(level1[m:m + block])
    assert halo_fraction(h, a, tile) <= halo_fraction(h, order_of(h, tile), tile) + 0.05


def test_coarsened_macro_blocks(monkeypatch):
    """Large systems grow the macro-blocks on a coarsened graph (16 consecutive sites per super-node): still a
    deterministic permutation with compact tiles, and every block but the ones moved to the end is made of whole
    super-nodes and starts on a tile boundary"""
    h = pb.graphene_rectangle(100.0, dtype=np.float32).hamiltonian.tocsr()    # 382 k sites >= 8 macro-blocks of 16 k
    n, tile, macro = h.shape[0], 256, 64
    monkeypatch.setenv("PBK_COARSE", "16")
    a = order_of(h, tile, macro)
    assert np.array_equal(np.sort(a), np.arange(n))
    assert np.array_equal(a, order_of(h, tile, macro))
    monkeypatch.setenv("PBK_COARSE", "1")
    fine = order_of(h, tile, macro)
    assert not np.array_equal(a, fine)                              # the coarse pass was really used
    assert halo_fraction(h, a, tile) < halo_fraction(h, fine, tile) + 0.03
    block = tile * macro
    first = a[:block]                                               # a regular block: whole super-nodes of 16 sites
    assert np.array_equal(np.unique(first // 16).repeat(16), np.sort(first) // 16)


def test_cubic_order_is_a_permutation():
    h = pb.cubic_anderson(20, disorder=1.0, dtype=np.float32).hamiltonian.tocsr()
    for macro in (0, 8):
        order = order_of(h, 256, macro)
        assert np.array_equal(np.sort(order), np.arange(h.shape[0]))
        assert halo_fraction(h, order, 256) < 2.0                   # natural order: 4.0 + 2 neighbours in the line


def test_dense_random_graph_falls_back_to_the_fine_macro_pass():
    """More than 16 distinct neighbouring super-nodes per super-node (no lattice structure in the caller's order): the
    coarse adjacency overflows and the macro-blocks are grown on the real graph; still a permutation"""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 140_000
    rows = np.repeat(np.arange(n), 4)
    cols = rng.integers(0, n, size=rows.size)
    a = sp.coo_matrix((np.ones(rows.size, np.float32), (rows, cols)), shape=(n, n)).tocsr()
    h = (a + a.T).tocsr()
    h.sort_indices()
    order = order_of(h, 256, 64)            # 8 macro-blocks of 16 k sites: the coarse pass is attempted first
    assert np.array_equal(np.sort(order), np.arange(n))
    assert np.array_equal(order, order_of(h, 256, 64))
