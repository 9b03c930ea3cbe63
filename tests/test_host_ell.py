"""Host side of the Hamiltonian hand-off (engine.cu: build_ell_host) against the oracle's restatement of
OptimizedHamiltonian::create_scaled / create_reordered (cppcore/src/kpm/OptimizedHamiltonian.cpp:55-152): same scale
factors, same per-element formulas in the Hamiltonian's scalar type (including the diagonal created by a non-zero `b`),
rows sorted by (new) column index, zero padding."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

import pybinding_b200 as pb
from oracle.oracle import OracleKPM
from pybinding_b200 import _lib


def host_ell(h, energy_range, order=None):
    lib = _lib.load()
    n = h.shape[0]
    indptr = np.ascontiguousarray(h.indptr, np.int32)
    indices = np.ascontiguousarray(h.indices, np.int32)
    data = np.ascontiguousarray(h.data)
    order_arr = None if order is None else np.ascontiguousarray(order, np.int32)
    optr = None if order is None else _lib.ptr(order_arr)
    k, pitch = ctypes.c_int32(), ctypes.c_int64()
    args = (_lib.DTYPES[np.dtype(h.dtype)], ctypes.c_int64(n), _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
            float(energy_range[0]), float(energy_range[1]), optr, ctypes.byref(k), ctypes.byref(pitch))
    assert lib.pbk_host_ell(*args, None, None) == 0
    val = np.zeros(k.value * pitch.value, h.dtype)
    col = np.zeros(k.value * pitch.value, np.int32)
    assert lib.pbk_host_ell(*args, _lib.ptr(val), _lib.ptr(col)) == 0
    return val.reshape(k.value, pitch.value), col.reshape(k.value, pitch.value), k.value, pitch.value


def ell_to_csr(val, col, n):
    k = val.shape[0]
    rows = np.tile(np.arange(n), k)
    m = sp.coo_matrix((val[:, :n].ravel(), (rows, col[:, :n].ravel())), shape=(n, n)).tocsr()   # padding adds zeros
    m.eliminate_zeros()
    m.sort_indices()
    return m


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("energy_range", [(-9.0, 9.0), (-7.5, 9.3)], ids=["b=0", "b!=0"])
def test_scaled_ell_equals_the_reference_restatement(dtype, energy_range):
    dtype = np.dtype(dtype)
    model = pb.graphene_rectangle(5.0, dtype=dtype, onsite=0.0 if energy_range[0] == -9.0 else 0.25,
                                  magnetic_field=300.0 if dtype.kind == "c" else 0.0)
    h = model.hamiltonian.tocsr()
    n = h.shape[0]
    # caller's order: create_scaled
    val, col, k, pitch = host_ell(h, energy_range)
    assert pitch % 32 == 0 and pitch >= n and np.all(col[:, n:] == 0) and np.all(val[:, n:] == 0)
    ref = OracleKPM(h, energy_range=energy_range, optimal_size=False, interleaved=False)   # AlgorithmConfig::reorder() == false
    info = ref.optimize_for([0], [0])
    expected = ref.optimized_matrix(info["nnz"])
    expected = (expected.real if dtype.kind != "c" else expected).astype(dtype)
    expected.eliminate_zeros()
    got = ell_to_csr(val, col, n)
    assert np.array_equal(got.indptr, expected.indptr) and np.array_equal(got.indices, expected.indices)
    assert np.array_equal(got.data, expected.data)                       # bit-identical scaling arithmetic
    # relabelled order: create_reordered with the reference's own breadth-first map
    src = n // 2
    ref2 = OracleKPM(h, energy_range=energy_range, optimal_size=True)
    info2 = ref2.optimize_for([src], [src])
    order = np.argsort(info2["reorder_map"]).astype(np.int32)             # order[new] = old
    val2, col2, k2, _ = host_ell(h, energy_range, order)
    expected2 = ref2.optimized_matrix(info2["nnz"])
    expected2 = (expected2.real if dtype.kind != "c" else expected2).astype(dtype)
    expected2.eliminate_zeros()
    got2 = ell_to_csr(val2, col2, n)
    assert np.array_equal(got2.indptr, expected2.indptr) and np.array_equal(got2.indices, expected2.indices)
    # `value * (2/a) - b * (2/a)` (OptimizedHamiltonian.cpp:124-127): the oracle is built with FMA contraction available
    # (-march=x86-64-v3), the engine's host code without, so the shifted diagonal may differ in the last bit
    eps = np.finfo(dtype).eps
    assert np.abs(got2.data - expected2.data).max() <= 2 * eps * np.abs(expected2.data).max()
    if energy_range[0] == -9.0:
        assert np.array_equal(got2.data, expected2.data)                 # b == 0: no shifted diagonal, bit-identical
    # every row is sorted by column and padded with (0, own row)
    for row in (0, 1, n // 2, n - 1):
        nnz_row = got2.indptr[row + 1] - got2.indptr[row]
        assert np.all(np.diff(col2[:nnz_row, row]) > 0)
        assert np.all(col2[nnz_row:, row] == row) and np.all(val2[nnz_row:, row] == 0)


def test_rejects_a_non_permutation():
    h = pb.graphene_rectangle(3.0, dtype=np.float32).hamiltonian.tocsr()
    lib = _lib.load()
    n = h.shape[0]
    order = np.zeros(n, np.int32)
    k, pitch = ctypes.c_int32(), ctypes.c_int64()
    rc = lib.pbk_host_ell(0, ctypes.c_int64(n), _lib.ptr(np.ascontiguousarray(h.indptr, np.int32)),
                          _lib.ptr(np.ascontiguousarray(h.indices, np.int32)), _lib.ptr(np.ascontiguousarray(h.data)),
                          -9.0, 9.0, _lib.ptr(order), ctypes.byref(k), ctypes.byref(pitch), None, None)
    assert rc == 1
