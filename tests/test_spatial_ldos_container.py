"""`SpatialLDOS` / `SiteSelection` (host-side containers of calc_spatial_ldos): the surface of the reference class
(pybinding/chebyshev.py:20-62 -- `structure_map(energy)`, `ldos(position, sublattice)`) without a GPU."""
import numpy as np

import pybinding_b200 as pb
from pybinding_b200.chebyshev import SiteSelection, SpatialLDOS


def test_spatial_ldos_container():
    model = pb.graphene_rectangle(4.0, dtype=np.float32)
    system = model.system
    contains = np.asarray(pb.Rectangle(2.0).contains(*system.positions))
    indices = np.flatnonzero(contains)
    sel = SiteSelection(system, indices)
    assert len(sel) == indices.size > 10
    energy = np.linspace(-1, 1, 11)
    data = np.arange(energy.size * len(sel), dtype=float).reshape(energy.size, len(sel))
    sl = SpatialLDOS(data, energy, sel)
    assert sl.data.shape == (energy.size, len(sl.structure))
    # structure map at the sampled energy closest to 0.33 (index 7: 0.4) carries one value per selected site
    smap = sl.structure_map(0.33)
    assert np.array_equal(smap.data, data[7]) and np.array_equal(smap.indices, indices) and np.array_equal(sl.ldos_at(0.33), data[7])
    # LDOS curve at the selected site nearest to a position == the column of that site
    pos = [0.3, -0.4]
    col = sel.find_nearest(pos)
    d2 = (sel.x - np.float32(pos[0])) ** 2 + (sel.y - np.float32(pos[1])) ** 2
    assert col == int(np.argmin(d2))
    series = sl.ldos(pos)
    assert np.array_equal(series.data, data[:, col]) and np.array_equal(series.variable, energy)
    # sublattice filter: the nearest selected B site
    b = sel.find_nearest(pos, "B")
    start, end = system.sublattice_range("B")
    assert start <= sel.indices[b] < end


def test_sublattice_range_without_the_method():
    """A real pybinding System has no `sublattice_range` in Python (cppmodule/src/system.cpp:84-94): the range is read
    off `system.sublattices`, an id array that compares with sublattice names (pybinding.support.alias.AliasArray)."""
    from pybinding_b200.chebyshev import sublattice_range

    class AliasIds(np.ndarray):
        names = {"A": 0, "B": 1}

        def __eq__(self, other):
            return np.asarray(self).__eq__(self.names[other] if isinstance(other, str) else other)

    class DuckSystem:   # the attributes chebyshev.py uses, nothing else
        def __init__(self, reference):
            self.positions = reference.positions
            self.num_sites = reference.num_sites
            start_b = reference.sublattice_range("B")[0]
            ids = np.zeros(self.num_sites, np.int8)
            ids[start_b:] = 1
            self.sublattices = ids.view(AliasIds)

    model = pb.graphene_rectangle(4.0, dtype=np.float32)
    duck = DuckSystem(model.system)
    assert not hasattr(duck, "sublattice_range")
    for name in ("", "A", "B"):
        assert sublattice_range(duck, name) == tuple(model.system.sublattice_range(name))
    sel = SiteSelection(duck, np.arange(duck.num_sites))
    b = sel.find_nearest([0.3, -0.4], "B")
    assert model.system.sublattice_range("B")[0] <= b
