"""Host-only checks of the synthetic model generators against the reference's own golden curves, through the oracle.

`lattice_model` (pybinding_b200/synthetic.py) reproduces `pb.Model(lattice, shape, ...)` for the models of the
reference's KPM and sweep tests that are not nearest-neighbour graphene rectangles:
 * `group6_tmd.monolayer_3band("MoS2")` + `pb.rectangle(6)`: 3 orbitals per site (tests/test_kpm.py:23-47, ldos[mos2].pbz)
 * `graphene.monolayer()` + `graphene.hexagon_ac(15)` [+ constant potentials] (tests/test_parallel.py:16-52)
"""
import os

import numpy as np
import pytest

import pybinding_b200 as pb
from pybinding_b200 import synthetic as syn
from oracle.oracle import OracleKPM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def parallel_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_parallel_baselines.npz"))


@pytest.mark.parametrize("kw", [dict(), dict(onsite=0.5), dict(magnetic_field=1e3)], ids=["pristine", "potential", "field"])
def test_generic_builder_equals_the_graphene_generator(kw):
    a = pb.graphene_rectangle(15, **kw)
    b = syn.graphene_monolayer(pb.Rectangle(15), **kw)
    assert a.hamiltonian.dtype == b.hamiltonian.dtype and a.hamiltonian.shape == b.hamiltonian.shape
    assert np.array_equal(a.hamiltonian.indptr, b.hamiltonian.indptr)
    assert np.array_equal(a.hamiltonian.indices, b.hamiltonian.indices)
    assert np.array_equal(a.hamiltonian.data, b.hamiltonian.data)
    assert np.array_equal(a.system.x, b.system.x) and np.array_equal(a.system.y, b.system.y)


def test_mos2_three_band_model_and_golden(golden):
    """tests/test_kpm.py:23-47 for "mos2": LDOS per orbital at [0, 0.07], Lorentz kernel, auto bounds"""
    model = syn.mos2_3band(pb.Rectangle(6))
    h = model.hamiltonian
    assert model.is_multiorbital and h.dtype == np.float32
    assert h.shape[0] == 3 * model.system.num_sites == model.system.hamiltonian_size
    assert np.diff(h.indptr).max() == 19            # onsite + 6 neighbours x 3 orbitals
    assert abs(h - h.T).max() == 0                  # real symmetric
    site = model.system.find_nearest([0, 0.07])
    idx = model.system.to_hamiltonian_indices(site)
    assert idx.tolist() == [3 * site, 3 * site + 1, 3 * site + 2]
    assert model.system.expanded_positions.x.size == h.shape[0]
    energy = np.linspace(0, 2, 25)
    ldos = OracleKPM(h, kernel="lorentz").calc_ldos(energy, 0.15, idx)
    assert ldos.shape == (25, 3)
    assert np.allclose(ldos, golden["ldos[mos2]"], rtol=1e-3, atol=1e-6)


def test_next_nearest_neighbour_graphene_has_ten_entries_per_row():
    model = syn.graphene_monolayer(pb.Rectangle(8), nearest_neighbors=2)
    h = model.hamiltonian
    assert np.diff(h.indptr).max() == 10            # 3 + 6 hoppings + the 3 t_nn onsite offset
    assert abs(h - h.T).max() == 0


def test_sweep_goldens_through_the_oracle(parallel_golden):
    """tests/test_parallel.py:16-52: LDOS at [0, 0] of a graphene hexagon against a constant potential"""
    energy = np.linspace(0, 0.1, 10)
    shape = syn.graphene_hexagon_ac(15)
    rows = []
    for v in np.linspace(0, 0.1, 10):
        model = syn.graphene_monolayer(shape, onsite=v)
        i = model.system.find_nearest([0, 0], "B")
        rows.append(OracleKPM(model.hamiltonian, kernel="lorentz").calc_ldos(energy, 0.15, [i])[:, 0])
    assert np.allclose(np.array(rows), parallel_golden["sweep.data"], rtol=1e-3, atol=1e-6)
    v1 = parallel_golden["ndsweep.variables.0"]
    v2 = parallel_golden["ndsweep.variables.1"]
    for a in (0, len(v1) - 1):          # two rows of the 5 x 4 grid are enough for the host-side check
        for b in range(len(v2)):
            model = syn.graphene_monolayer(shape, onsite=float(np.float32(v1[a]) + np.float32(v2[b])))
            i = model.system.find_nearest([0, 0])
            ldos = OracleKPM(model.hamiltonian, kernel="lorentz").calc_ldos(energy, 0.15, [i])[:, 0]
            assert np.allclose(ldos, parallel_golden["ndsweep.data"][a, b], rtol=1e-3, atol=1e-6)
