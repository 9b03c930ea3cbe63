"""Synthetic tight-binding models for the KPM hot path (inputs only, not a model builder).

pybinding's model builder (Lattice / Shape / Foundation / modifiers) is out of scope for this
repository; the KPM engine only consumes what crosses the boundary: a CSR Hamiltonian
(`model.hamiltonian`) and site positions (`model.system`).  This module produces exactly those
two things for the lattices named in BASELINE.json, in the *same site order* the reference would
produce, so that the stochastic starters (which are drawn per site index) and the reference's
golden curves can be reproduced:

* site order: sublattice-major, then lattice vector a2, then a1 fastest
  (reference: cppcore/src/system/Foundation.cpp:24-48, 127-164)
* bounding box in lattice coordinates with +/-1 padding (Foundation.cpp:6-22)
* polygon containment by ray casting in float32 (cppcore/src/system/Shape.cpp:52-90)
* iterative removal of sites with fewer than `min_neighbors` neighbours (Foundation.cpp:51-101;
  graphene sets min_neighbors=2: pybinding/repository/graphene/lattice.py:65)
* graphene constants a=0.24595, a_cc=0.142, t=-2.8 (pybinding/repository/graphene/constants.py:3-5)
* Peierls phase of `constant_magnetic_field` (pybinding/repository/graphene/modifiers.py:54-73)
* zero onsite terms are not stored in the CSR matrix (HamiltonianModifiers.hpp:107)

The objects returned duck-type the parts of `pb.Model` / `pb.System` which the KPM facade uses
(reference: cppcore/src/KPM.cpp:55-148, cppmodule/src/system.cpp:85-94).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

__all__ = ["Model", "System", "Positions", "Rectangle", "Polygon", "graphene_rectangle", "cubic_anderson",
           "lattice_model", "graphene_monolayer", "graphene_hexagon_ac", "mos2_3band",
           "GRAPHENE_A", "GRAPHENE_ACC", "GRAPHENE_T", "GRAPHENE_T_NN"]

GRAPHENE_A = 0.24595    # [nm] unit cell length
GRAPHENE_ACC = 0.142    # [nm] carbon-carbon distance
GRAPHENE_T = -2.8       # [eV] nearest neighbour hopping
GRAPHENE_T_NN = 0.1     # [eV] next-nearest neighbour hopping (pybinding/repository/graphene/constants.py:6)
_HBAR = 6.58211899e-16  # [eV*s]  (pybinding/constants.py)
_PHI0 = 2 * math.pi * _HBAR


@dataclass
class Positions:
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __getitem__(self, i):
        return (self.x, self.y, self.z)[i]


class Rectangle:
    """Minimal stand-in for `pb.rectangle` (only `contains` is used, by calc_spatial_ldos)"""

    def __init__(self, x, y=None, center=(0.0, 0.0)):
        y = y or x
        self.x0, self.y0 = x / 2, y / 2
        self.center = center
        cx, cy = center
        self.vertices = [(cx + self.x0, cy + self.y0), (cx + self.x0, cy - self.y0),
                         (cx - self.x0, cy - self.y0), (cx - self.x0, cy + self.y0)]

    def contains(self, x, y, z=None):
        return _within_polygon(np.asarray(x, np.float32), np.asarray(y, np.float32), self.vertices)


class Polygon:
    """Minimal stand-in for `pb.Polygon` (pybinding/shape.py): vertices + `lattice_offset` + `contains`"""

    def __init__(self, vertices, lattice_offset=(0.0, 0.0)):
        self.vertices = [tuple(float(c) for c in v[:2]) for v in vertices]
        self.lattice_offset = tuple(lattice_offset)

    def contains(self, x, y, z=None):
        return _within_polygon(np.asarray(x, np.float32), np.asarray(y, np.float32), self.vertices)


class System:
    """Positions + sublattice bookkeeping.  `orbitals[i]` = number of orbitals of sublattice i (default 1): the
    Hamiltonian rows of a site are consecutive, sublattice blocks follow each other
    (System::to_hamiltonian_indices / hamiltonian_size / expanded_positions, cppcore/src/system/System.cpp:8-97)"""

    def __init__(self, x, y, z, sub_starts, sub_names, orbitals=None):
        self.positions = Positions(np.ascontiguousarray(x, np.float32),
                                   np.ascontiguousarray(y, np.float32),
                                   np.ascontiguousarray(z, np.float32))
        self._sub_starts = list(sub_starts)  # len = nsub + 1
        self._sub_names = list(sub_names)
        self._orbitals = list(orbitals) if orbitals is not None else [1] * len(self._sub_names)
        self._ham_starts = [0]
        for i, norb in enumerate(self._orbitals):
            self._ham_starts.append(self._ham_starts[-1] + norb * (self._sub_starts[i + 1] - self._sub_starts[i]))

    @property
    def num_sites(self):
        return int(self.positions.x.size)

    @property
    def hamiltonian_size(self):
        return int(self._ham_starts[-1])

    @property
    def is_multiorbital(self):
        return any(norb > 1 for norb in self._orbitals)

    @property
    def expanded_positions(self):
        if not self.is_multiorbital:
            return self.positions
        reps = np.concatenate([np.full(self._sub_starts[i + 1] - self._sub_starts[i], norb, np.int64)
                               for i, norb in enumerate(self._orbitals)])
        return Positions(np.repeat(self.positions.x, reps), np.repeat(self.positions.y, reps),
                         np.repeat(self.positions.z, reps))

    @property
    def x(self):
        return self.positions.x

    @property
    def y(self):
        return self.positions.y

    @property
    def z(self):
        return self.positions.z

    def sublattice_range(self, sublattice=""):
        if not sublattice:
            return 0, self.num_sites
        if sublattice not in self._sub_names:
            raise IndexError("There is no sublattice named '{}'".format(sublattice))
        i = self._sub_names.index(sublattice)
        return self._sub_starts[i], self._sub_starts[i + 1]

    def find_nearest(self, position, sublattice=""):
        """First site with the smallest float32 distance (reference: System.cpp:68-82)"""
        start, end = self.sublattice_range(sublattice)
        p = np.zeros(3, np.float32)
        p[:len(position)] = np.asarray(position, np.float32)
        dx = self.positions.x[start:end] - p[0]
        dy = self.positions.y[start:end] - p[1]
        dz = self.positions.z[start:end] - p[2]
        d = np.sqrt(dx * dx + dy * dy + dz * dz)
        return int(start + np.argmin(d))

    def to_hamiltonian_indices(self, system_index):
        for i, norb in enumerate(self._orbitals):
            if self._sub_starts[i] <= system_index < self._sub_starts[i + 1]:
                first = self._ham_starts[i] + (system_index - self._sub_starts[i]) * norb
                return np.arange(first, first + norb, dtype=np.int32)
        raise IndexError("to_hamiltonian_indices: site index out of range")


@dataclass
class Model:
    hamiltonian: sp.csr_matrix
    system: System
    description: str = ""
    meta: dict = field(default_factory=dict)

    def eval(self):
        return self

    @property
    def is_complex(self):
        return np.iscomplexobj(self.hamiltonian.data)

    @property
    def is_double(self):
        return self.hamiltonian.dtype in (np.float64, np.complex128)

    @property
    def is_multiorbital(self):
        return bool(getattr(self.system, "is_multiorbital", False))

    @property
    def raw_hamiltonian(self):
        return self.hamiltonian


def _within_polygon(px, py, vertices):
    """Ray casting in float32, side by side as in Shape.cpp:52-90"""
    vx = np.array([v[0] for v in vertices], np.float32)
    vy = np.array([v[1] for v in vertices], np.float32)
    inside = np.zeros(px.shape, dtype=bool)
    n = len(vertices)
    j = n - 1
    for i in range(n):
        x1, x2, y1, y2 = vx[i], vx[j], vy[i], vy[j]
        j = i
        diff = abs(np.float32(y1 - y2))
        scale = abs(np.float32(y1 + y2))
        eps = np.finfo(np.float32).eps
        if diff <= eps * scale or diff <= np.finfo(np.float32).tiny:
            continue  # ray parallel to this side
        k = np.float32((x2 - x1) / (y2 - y1))
        intersects_y = (y1 > py) != (y2 > py)
        x_side = k * (py - y1) + x1
        intersects_x = px > x_side
        inside ^= (intersects_y & intersects_x)
    return inside


def _resolve_dtype(dtype, is_complex, is_double):
    if dtype is not None:
        return np.dtype(dtype)
    return np.dtype({(False, False): np.float32, (True, False): np.complex64,
                     (False, True): np.float64, (True, True): np.complex128}[(is_complex, is_double)])


def _rows_to_csr(n, cols, vals, dtype):
    """Build a sorted CSR matrix from per-row candidate lists (col < 0 marks an empty slot)"""
    order = np.argsort(np.where(cols < 0, np.iinfo(np.int64).max, cols), axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1)
    vals = np.take_along_axis(vals, order, axis=1)
    mask = cols >= 0
    counts = mask.sum(axis=1)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    if indptr[-1] >= 2**31:
        raise ValueError("Hamiltonian has too many non-zeros for int32 indices")
    h = sp.csr_matrix((vals[mask].astype(dtype, copy=False), cols[mask].astype(np.int32),
                       indptr.astype(np.int32)), shape=(n, n))
    h.has_sorted_indices = True
    return h


def graphene_rectangle(width, height=None, *, onsite=0.0, disorder=0.0, disorder_seed=0,
                       magnetic_field=0.0, dtype=None, min_neighbors=2, t=GRAPHENE_T):
    """graphene.monolayer() + pb.rectangle(width, height) [+ onsite / Peierls modifiers]

    Parameters
    ----------
    width, height : float
        Rectangle size in nm (`height` defaults to `width`, like `pb.rectangle`).
    onsite : float
        `pb.constant_potential(onsite)`; zero means no stored diagonal.
    disorder : float
        Uniform onsite disorder U(-disorder/2, disorder/2) from `numpy.random.default_rng(seed)`.
    magnetic_field : float
        `graphene.constant_magnetic_field(B)` in Tesla; makes the Hamiltonian complex.
    dtype : numpy dtype, optional
        Force the scalar type (default: float32, or complex64 with a magnetic field).
    min_neighbors : int
        2 for `pybinding.repository.graphene` (Python), 1 for the C++ test fixture lattice.
    """
    height = height or width
    f32 = np.float32
    a, acc = f32(GRAPHENE_A), f32(GRAPHENE_ACC)
    a1 = np.array([a, 0], f32)
    a2 = np.array([a / f32(2), a / f32(2) * f32(math.sqrt(3.0))], f32)
    sub_offsets = [np.array([0, -acc / f32(2)], f32), np.array([0, acc / f32(2)], f32)]

    x0, y0 = width / 2, height / 2
    vertices = [(x0, y0), (x0, -y0), (-x0, -y0), (-x0, y0)]

    # Foundation bounds (Foundation.cpp:6-22): lattice coordinates of the vertices, truncated, +/-1
    lat = np.array([[a1[0], a2[0]], [a1[1], a2[1]]], np.float64)
    v = np.array([np.linalg.solve(lat, np.array(p, np.float64)).astype(f32) for p in vertices])
    v = np.trunc(v).astype(np.int64)
    lower = v.min(axis=0) - 1
    upper = v.max(axis=0) + 1
    size0, size1 = (upper - lower + 1).tolist()
    block = size0 * size1

    # Positions (generate_positions, Foundation.cpp:24-48), float32 with the reference's op order
    origin = f32(lower[0]) * a1 + f32(lower[1]) * a2
    ia = np.arange(size0, dtype=f32)
    ib = np.arange(size1, dtype=f32)
    px = np.empty((2, size1, size0), f32)
    py = np.empty((2, size1, size0), f32)
    for n in range(2):
        ps = origin + sub_offsets[n]
        pbx = np.where(ib == 0, ps[0], ps[0] + ib * a2[0]).astype(f32)
        pby = np.where(ib == 0, ps[1], ps[1] + ib * a2[1]).astype(f32)
        px[n] = pbx[:, None] + ia[None, :] * a1[0]
        py[n] = pby[:, None] + ia[None, :] * a1[1]

    valid = _within_polygon(px.ravel(), py.ravel(), vertices).reshape(2, size1, size0)

    # Neighbours: A(i,j) - B(i,j), B(i+1,j-1), B(i,j-1)  (graphene/lattice.py:41-45)
    def shifted(arr, da, db, fill):
        """arr[b+db, a+da] with out-of-range filled"""
        out = np.full(arr.shape, fill, arr.dtype)
        sb = slice(max(0, -db), arr.shape[0] - max(0, db))
        sa = slice(max(0, -da), arr.shape[1] - max(0, da))
        tb = slice(max(0, db), arr.shape[0] - max(0, -db))
        ta = slice(max(0, da), arr.shape[1] - max(0, -da))
        out[sb, sa] = arr[tb, ta]
        return out

    hops = [(0, 0), (1, -1), (0, -1)]  # (da, db) from A to B
    while True:  # remove_dangling fix point (Foundation.cpp:51-101)
        va, vb = valid[0], valid[1]
        cnt_a = sum(shifted(vb, da, db, False).astype(np.int8) for da, db in hops)
        cnt_b = sum(shifted(va, -da, -db, False).astype(np.int8) for da, db in hops)
        new_a = va & (cnt_a >= min_neighbors)
        new_b = vb & (cnt_b >= min_neighbors)
        if new_a.sum() == va.sum() and new_b.sum() == vb.sum():
            break
        valid = np.stack([new_a, new_b])

    # Final indices: all valid A sites then all valid B sites (get_finalized_indices, :127-164)
    flat_valid = valid.ravel()
    index = np.full(2 * block, -1, np.int64)
    n_sites = int(flat_valid.sum())
    index[flat_valid] = np.arange(n_sites)
    index = index.reshape(2, size1, size0)
    n_a = int(valid[0].sum())

    x = px.ravel()[flat_valid]
    y = py.ravel()[flat_valid]
    z = np.zeros(n_sites, f32)

    is_complex = magnetic_field != 0
    dtype = _resolve_dtype(dtype, is_complex, False)
    if is_complex and dtype.kind != "c":
        raise ValueError("a magnetic field requires a complex dtype")
    real_dtype = np.dtype(np.float32 if dtype in (np.float32, np.complex64) else np.float64)

    has_onsite = (onsite != 0) or (disorder != 0)
    ncand = 4 if has_onsite else 3
    cols = np.full((n_sites, ncand), -1, np.int64)
    vals = np.zeros((n_sites, ncand), dtype)

    ia_idx = index[0][valid[0]]  # final indices of A sites, in order
    ib_idx = index[1][valid[1]]
    for s, (da, db) in enumerate(hops):
        nb_of_a = shifted(index[1], da, db, -1)[valid[0]]      # B neighbour of each A site
        na_of_b = shifted(index[0], -da, -db, -1)[valid[1]]    # A neighbour of each B site
        cols[ia_idx, s] = nb_of_a
        cols[ib_idx, s] = na_of_b

    hop = np.full((n_sites, 3), t, dtype)
    if is_complex:
        # energy * exp(1j * const * 0.5*B*(y1+y2) * (x1-x2)); the B->A entry is the conjugate
        const = real_dtype.type(1e-18 * 2 * math.pi / _PHI0)
        for s in range(3):
            c = cols[:n_a, s]
            ok = c >= 0
            x1 = x[:n_a][ok].astype(real_dtype)
            y1 = y[:n_a][ok].astype(real_dtype)
            x2 = x[c[ok]].astype(real_dtype)
            y2 = y[c[ok]].astype(real_dtype)
            peierls = (real_dtype.type(0.5 * magnetic_field) * (y1 + y2)) * (x1 - x2)
            phase = np.exp(1j * (const * peierls)).astype(dtype)
            va_ = np.full(n_a, t, dtype)
            va_[ok] = (dtype.type(t) * phase)
            hop[:n_a, s] = va_
            # conjugate entries on the B rows: find them through the same (da, db) relation
            rows_b = c[ok]
            hop[rows_b, s] = np.conj(va_[ok])
    vals[:, :3] = np.where(cols[:, :3] >= 0, hop, 0)

    if has_onsite:
        diag = np.full(n_sites, onsite, np.float64)
        if disorder != 0:
            rng = np.random.default_rng(disorder_seed)
            diag = diag + rng.uniform(-disorder / 2, disorder / 2, n_sites)
        nz = diag != 0
        cols[nz, 3] = np.arange(n_sites)[nz]
        vals[:, 3] = diag.astype(dtype)

    h = _rows_to_csr(n_sites, cols, vals, dtype)
    system = System(x, y, z, [0, n_a, n_sites], ["A", "B"])
    desc = "graphene.monolayer() {}x{} nm rectangle".format(width, height)
    return Model(h, system, desc, dict(width=width, height=height, onsite=onsite, disorder=disorder,
                                       magnetic_field=magnetic_field, min_neighbors=min_neighbors,
                                       volume=float(width) * float(height)))


def cubic_anderson(length, *, disorder=4.0, seed=0, t=-1.0, periodic=True, dtype=np.float32):
    """Simple-cubic Anderson lattice `length`^3: hopping `t` to 6 neighbours, onsite U(-W/2, W/2)

    Site index = (z*L + y)*L + x; lattice constant 1 nm.
    """
    L = int(length)
    n = L ** 3
    dtype = np.dtype(dtype)
    idx = np.arange(n, dtype=np.int64).reshape(L, L, L)  # [z, y, x]
    cols = np.full((n, 7), -1, np.int64)
    vals = np.zeros((n, 7), dtype)
    s = 0
    for axis in range(3):
        for shift in (-1, 1):
            nb = np.roll(idx, -shift, axis=axis)
            if not periodic or L <= 2:
                sl = [slice(None)] * 3
                sl[axis] = slice(L - 1, L) if shift == 1 else slice(0, 1)
                nb = nb.copy()
                nb[tuple(sl)] = -1
            cols[:, s] = nb.ravel()
            vals[:, s] = np.where(nb.ravel() >= 0, t, 0)
            s += 1
    rng = np.random.default_rng(seed)
    diag = rng.uniform(-disorder / 2, disorder / 2, n) if disorder != 0 else np.zeros(n)
    nz = diag != 0
    cols[nz, 6] = np.arange(n)[nz]
    vals[:, 6] = diag.astype(dtype)
    h = _rows_to_csr(n, cols, vals, dtype)

    zz, yy, xx = np.meshgrid(np.arange(L, dtype=np.float32), np.arange(L, dtype=np.float32),
                             np.arange(L, dtype=np.float32), indexing="ij")
    system = System(xx.ravel(), yy.ravel(), zz.ravel(), [0, n], ["A"])
    return Model(h, system, "simple cubic Anderson {}^3, W={}".format(L, disorder),
                 dict(length=L, disorder=disorder, periodic=periodic, volume=float(n)))


# ------------------------------------------------------------------------------------------------
# Generic 2-D lattice + polygon builder: the models of the reference's own KPM / sweep tests that are not
# nearest-neighbour graphene rectangles (tests/test_kpm.py:23 `group6_tmd.monolayer_3band("MoS2")`,
# tests/test_parallel.py:16-52 `graphene.hexagon_ac`) and lattices with other ELL widths (next-nearest-neighbour
# graphene).  Same conventions as `graphene_rectangle` above, written for clarity rather than for 38 M sites.
# ------------------------------------------------------------------------------------------------
def _as_matrix(energy, rows, cols=None):
    """Scalar / 1-D list (diagonal) / 2-D list -> matrix, like Lattice.add_one_sublattice / add_one_hopping"""
    e = np.atleast_1d(np.asarray(energy))
    if e.ndim == 1:
        if e.size == 1 and rows == 1:
            return e.reshape(1, 1)
        return np.diag(e)
    return e


def lattice_model(a1, a2, sublattices, hoppings, shape, *, min_neighbors=1, onsite=0.0, magnetic_field=0.0,
                  dtype=None, description=""):
    """`pb.Model(lattice, shape [, pb.constant_potential(onsite)] [, constant_magnetic_field(B)])` as CSR + positions

    Parameters
    ----------
    a1, a2 : (x, y) primitive vectors [nm]
    sublattices : list of (name, (x, y) offset, onsite energy: scalar, list (diagonal) or matrix)
    hoppings : list of ((da, db), from_name, to_name, energy: scalar or matrix with rows = orbitals of `from`)
        One entry per `Lattice.add_hoppings` term; the conjugate is added automatically
        (Hamiltonian.hpp:85-88: H(i, j) = t, H(j, i) = conj(t) with j = the site at i + (da, db)).
    shape : object with `.vertices` and optional `.lattice_offset` (`Polygon`, `Rectangle`)
    """
    f32 = np.float32
    a1 = np.array(a1, f32)
    a2 = np.array(a2, f32)
    names = [sub[0] for sub in sublattices]
    norbs = [int(_as_matrix(sub[2], 1).shape[0]) if np.ndim(sub[2]) > 0 else 1 for sub in sublattices]
    order = sorted(range(len(names)), key=lambda i: norbs[i])      # OptimizedUnitCell: stable sort by orbital count
    sublattices = [sublattices[i] for i in order]
    names = [names[i] for i in order]
    norbs = [norbs[i] for i in order]
    nsub = len(names)
    vertices = [tuple(v[:2]) for v in shape.vertices]
    offset = np.array(getattr(shape, "lattice_offset", (0.0, 0.0))[:2], f32)

    lat = np.array([[a1[0], a2[0]], [a1[1], a2[1]]], np.float64)
    v = np.array([np.linalg.solve(lat, np.array(p, np.float64)).astype(f32) for p in vertices])
    v = np.trunc(v).astype(np.int64)
    lower = v.min(axis=0) - 1
    upper = v.max(axis=0) + 1
    size0, size1 = (upper - lower + 1).tolist()

    origin = offset + f32(lower[0]) * a1 + f32(lower[1]) * a2      # Lattice::calc_position(bounds.first)
    ia = np.arange(size0, dtype=f32)
    ib = np.arange(size1, dtype=f32)
    px = np.empty((nsub, size1, size0), f32)
    py = np.empty((nsub, size1, size0), f32)
    for n in range(nsub):
        ps = origin + np.array(sublattices[n][1][:2], f32)
        pbx = np.where(ib == 0, ps[0], ps[0] + ib * a2[0]).astype(f32)
        pby = np.where(ib == 0, ps[1], ps[1] + ib * a2[1]).astype(f32)
        px[n] = pbx[:, None] + ia[None, :] * a1[0]
        py[n] = pby[:, None] + ia[None, :] * a1[1]
    valid = _within_polygon(px.ravel(), py.ravel(), vertices).reshape(nsub, size1, size0)

    def shifted(arr, da, db, fill):
        out = np.full(arr.shape, fill, arr.dtype)
        sb = slice(max(0, -db), arr.shape[0] - max(0, db))
        sa = slice(max(0, -da), arr.shape[1] - max(0, da))
        tb = slice(max(0, db), arr.shape[0] - max(0, -db))
        ta = slice(max(0, da), arr.shape[1] - max(0, -da))
        out[sb, sa] = arr[tb, ta]
        return out

    terms = [((int(rel[0]), int(rel[1])), names.index(fr), names.index(to), en) for rel, fr, to, en in hoppings]
    while True:  # remove_dangling fix point (Foundation.cpp:51-101): hoppings and their conjugates both count
        counts = [np.zeros((size1, size0), np.int32) for _ in range(nsub)]
        for (da, db), fr, to, _ in terms:
            counts[fr] += shifted(valid[to], da, db, False)
            counts[to] += shifted(valid[fr], -da, -db, False)
        new = np.stack([valid[n] & (counts[n] >= min_neighbors) for n in range(nsub)])
        if new.sum() == valid.sum():
            break
        valid = new

    flat_valid = valid.ravel()
    n_sites = int(flat_valid.sum())
    index = np.full(flat_valid.size, -1, np.int64)
    index[flat_valid] = np.arange(n_sites)
    index = index.reshape(nsub, size1, size0)
    sub_starts = [0]
    for n in range(nsub):
        sub_starts.append(sub_starts[-1] + int(valid[n].sum()))
    ham_starts = [0]
    for n in range(nsub):
        ham_starts.append(ham_starts[-1] + norbs[n] * (sub_starts[n + 1] - sub_starts[n]))
    x = px.ravel()[flat_valid]
    y = py.ravel()[flat_valid]

    is_complex = magnetic_field != 0 or any(np.iscomplexobj(np.asarray(t[3])) for t in terms) \
        or any(np.iscomplexobj(np.asarray(sub[2])) for sub in sublattices)
    dtype = _resolve_dtype(dtype, is_complex, False)
    if is_complex and dtype.kind != "c":
        raise ValueError("complex terms require a complex dtype")
    real_dtype = np.dtype(np.float32 if dtype in (np.float32, np.complex64) else np.float64)

    def ham_index(sub, site, orbital):
        return ham_starts[sub] + (site - sub_starts[sub]) * norbs[sub] + orbital

    rows, cols, vals = [], [], []
    for n in range(nsub):   # onsite terms; zeros are not stored (HamiltonianModifiers.hpp:149)
        energy = _as_matrix(sublattices[n][2], norbs[n]).astype(dtype) + np.eye(norbs[n], dtype=dtype) * dtype.type(onsite)
        sites = np.arange(sub_starts[n], sub_starts[n + 1])
        for p in range(norbs[n]):
            for q in range(norbs[n]):
                if energy[p, q] != 0:
                    rows.append(ham_index(n, sites, p)); cols.append(ham_index(n, sites, q))
                    vals.append(np.full(sites.size, energy[p, q], dtype))
    const = real_dtype.type(1e-18 * 2 * math.pi / _PHI0)
    for (da, db), fr, to, energy in terms:
        src = index[fr]
        dst = shifted(index[to], da, db, -1)
        ok = (src >= 0) & (dst >= 0)
        i_site, j_site = src[ok], dst[ok]
        h = _as_matrix(energy, norbs[fr], norbs[to]).astype(dtype)
        phase = None
        if magnetic_field != 0:
            x1, y1 = x[i_site].astype(real_dtype), y[i_site].astype(real_dtype)
            x2, y2 = x[j_site].astype(real_dtype), y[j_site].astype(real_dtype)
            peierls = (real_dtype.type(0.5 * magnetic_field) * (y1 + y2)) * (x1 - x2)
            phase = np.exp(1j * (const * peierls)).astype(dtype)
        for p in range(norbs[fr]):
            for q in range(norbs[to]):
                if h[p, q] == 0:
                    continue
                value = np.full(i_site.size, h[p, q], dtype) if phase is None else (h[p, q] * phase).astype(dtype)
                r, c = ham_index(fr, i_site, p), ham_index(to, j_site, q)
                rows += [r, c]; cols += [c, r]; vals += [value, np.conj(value)]
    size = ham_starts[-1]
    coo = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(size, size))
    h = coo.tocsr()
    h.sort_indices()
    h = sp.csr_matrix((h.data.astype(dtype), h.indices.astype(np.int32), h.indptr.astype(np.int32)), shape=(size, size))
    h.has_sorted_indices = True
    system = System(x, y, np.zeros(n_sites, f32), sub_starts, names, norbs)
    return Model(h, system, description, dict(onsite=onsite, magnetic_field=magnetic_field, min_neighbors=min_neighbors))


def graphene_monolayer(shape, *, nearest_neighbors=1, onsite=0.0, magnetic_field=0.0, dtype=None, t=GRAPHENE_T,
                       t_nn=GRAPHENE_T_NN):
    """`graphene.monolayer(nearest_neighbors)` (pybinding/repository/graphene/lattice.py:6-66) on any polygon

    nearest_neighbors = 2 adds the six second-neighbour hoppings and the 3 t_nn onsite offset: 3 + 6 + 1 = 10 entries
    per row, a lattice outside the specialised ELL widths of the step kernels."""
    a, acc = GRAPHENE_A, GRAPHENE_ACC
    offset = 0.0 if nearest_neighbors < 2 else 3 * t_nn
    subs = [("A", (0.0, -acc / 2), offset), ("B", (0.0, acc / 2), offset)]
    hops = [((0, 0), "A", "B", t), ((1, -1), "A", "B", t), ((0, -1), "A", "B", t)]
    if nearest_neighbors >= 2:
        for rel in ((0, -1), (1, -1), (1, 0)):
            hops += [(rel, "A", "A", t_nn), (rel, "B", "B", t_nn)]
    if nearest_neighbors >= 3:
        raise ValueError("nearest_neighbors > 2 is not provided")
    return lattice_model([a, 0.0], [a / 2, a / 2 * math.sqrt(3.0)], subs, hops, shape, min_neighbors=2, onsite=onsite,
                         magnetic_field=magnetic_field, dtype=dtype,
                         description="graphene.monolayer(nearest_neighbors={})".format(nearest_neighbors))


def graphene_hexagon_ac(side_width, lattice_offset=(-GRAPHENE_A / 2, 0.0)):
    """`graphene.hexagon_ac(side_width)` (pybinding/repository/graphene/shape.py:8-34): armchair edges on all sides"""
    side_atoms = math.ceil((side_width / GRAPHENE_ACC + 1) * 2 / 3)
    side_atoms += side_atoms % 2
    side_width = (3 / 2 * side_atoms - 1) * GRAPHENE_ACC - GRAPHENE_ACC / 2
    x0 = side_width * math.sqrt(3) / 2
    y0 = side_width
    return Polygon([(0, y0), (x0, y0 / 2), (x0, -y0 / 2), (0, -y0), (-x0, -y0 / 2), (-x0, y0 / 2)], lattice_offset)


def mos2_3band(shape, name="MoS2", dtype=None):
    """`group6_tmd.monolayer_3band(name)` (pybinding/repository/group6_tmd.py:19-113): one metal sublattice with three
    orbitals on a triangular lattice, 3 x 3 hopping matrices to the six neighbours (ELL width 21)"""
    params = {"MoS2": [0.3190, 1.046, 2.104, -0.184, 0.401, 0.507, 0.218, 0.338, 0.057],
              "WS2": [0.3191, 1.130, 2.275, -0.206, 0.567, 0.536, 0.286, 0.384, -0.061]}
    a, eps1, eps2, t0, t1, t2, t11, t12, t22 = params[name]
    rt3 = math.sqrt(3)
    h1 = [[t0, -t1, t2],
          [t1, t11, -t12],
          [t2, t12, t22]]
    h2 = [[t0, 1 / 2 * t1 + rt3 / 2 * t2, rt3 / 2 * t1 - 1 / 2 * t2],
          [-1 / 2 * t1 + rt3 / 2 * t2, 1 / 4 * t11 + 3 / 4 * t22, rt3 / 4 * (t11 - t22) - t12],
          [-rt3 / 2 * t1 - 1 / 2 * t2, rt3 / 4 * (t11 - t22) + t12, 3 / 4 * t11 + 1 / 4 * t22]]
    h3 = [[t0, -1 / 2 * t1 - rt3 / 2 * t2, rt3 / 2 * t1 - 1 / 2 * t2],
          [1 / 2 * t1 - rt3 / 2 * t2, 1 / 4 * t11 + 3 / 4 * t22, rt3 / 4 * (t22 - t11) + t12],
          [-rt3 / 2 * t1 - 1 / 2 * t2, rt3 / 4 * (t22 - t11) - t12, 3 / 4 * t11 + 1 / 4 * t22]]
    metal = name[:2] if name[1].islower() else name[:1]
    subs = [(metal, (0.0, 0.0), [eps1, eps2, eps2])]
    hops = [((1, 0), metal, metal, h1), ((0, -1), metal, metal, h2), ((1, -1), metal, metal, h3)]
    return lattice_model([a, 0.0], [a / 2, rt3 / 2 * a], subs, hops, shape, min_neighbors=1, dtype=dtype,
                         description="group6_tmd.monolayer_3band({})".format(name))
