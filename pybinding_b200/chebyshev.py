"""Computations based on Chebyshev polynomial expansion -- B200 engine behind pybinding's KPM API

Drop-in for `pybinding.chebyshev` (reference: pybinding/chebyshev.py:65-445): the same `kpm()`
factory, `KPM` class, kernel factories and result types.  `KPM` forwards to an `impl` object; here the
impl is `_CudaImpl`, the Python counterpart of the reference's C++ facade `cpb::KPM`
(cppcore/src/KPM.cpp:7-148) + its pybind11 binding (cppmodule/src/kpm.cpp:76-102), which calls the
GPU engine through the C ABI of libpbkpm.so (include/pbkpm.h).  There is no CPU path in this module.
"""
import ctypes as C
import threading
import warnings

import numpy as np

from . import _lib
from . import results

__all__ = ['KPM', 'kpm', 'kpm_cuda', 'SpatialLDOS', 'SiteSelection', 'Deferred',
           'jackson_kernel', 'lorentz_kernel', 'dirichlet_kernel', 'sublattice_range']


def sublattice_range(system, sublattice=""):
    """`System::sublattice_range` (cppcore/src/system/System.cpp:50-66): the contiguous block of sites of one sublattice

    The reference does not expose this method to Python (cppmodule/src/system.cpp:84-94), so for a real `pb.System`
    the block is read off `system.sublattices` (site IDs, sublattice-major order; an `AliasArray` that compares with
    names).  Systems which do provide `sublattice_range` (pybinding_b200.synthetic.System) are asked directly.
    """
    if hasattr(system, "sublattice_range"):
        start, end = system.sublattice_range(sublattice)
        return int(start), int(end)
    num_sites = int(getattr(system, "num_sites", len(np.asarray(system.positions[0]))))
    if not sublattice:
        return 0, num_sites
    ids = np.asarray(system.sublattices == sublattice)
    if ids.shape == ():   # a plain id array compared with a name: look the name up in the lattice
        names = getattr(getattr(system, "lattice", None), "sublattices", {})
        if sublattice not in names:
            raise IndexError("There is no sublattice named '{}'".format(sublattice))
        ids = np.asarray(system.sublattices) == getattr(names[sublattice], "alias_id", names[sublattice])
    hits = np.flatnonzero(ids)
    if hits.size == 0:
        raise IndexError("There is no sublattice named '{}'".format(sublattice))
    return int(hits[0]), int(hits[-1]) + 1


class SiteSelection:
    """The sites a spatial LDOS was computed for: the role `system[shape.contains(...)]` (a sliced `System` /
    `StructureMap`) plays in the reference (pybinding/chebyshev.py:223-227) for models that are not pybinding systems.
    Holds the Hamiltonian indices and positions of the selected sites."""

    def __init__(self, system, indices, data=None):
        self.indices = np.asarray(indices, np.int64)
        x, y, z = system.positions
        self.x, self.y, self.z = (np.asarray(a)[self.indices] for a in (x, y, z))
        self._system = system
        self.data = data

    def __len__(self):
        return int(self.indices.size)

    def find_nearest(self, position, sublattice=""):
        """Position (column of the LDOS table) of the selected site closest to `position`"""
        keep = np.arange(len(self))
        if sublattice:
            start, end = sublattice_range(self._system, sublattice)
            keep = keep[(self.indices >= start) & (self.indices < end)]
            if keep.size == 0:
                raise IndexError("no selected site on sublattice '{}'".format(sublattice))
        p = np.zeros(3, np.float32)
        p[:len(position)] = np.asarray(position, np.float32)
        d = (self.x[keep] - p[0]) ** 2 + (self.y[keep] - p[1]) ** 2 + (self.z[keep] - p[2]) ** 2
        return int(keep[np.argmin(d)])

    def with_data(self, data):
        """A copy carrying one value per selected site (what `StructureMap.with_data` returns in the reference)"""
        return SiteSelection(self._system, self.indices, np.asarray(data))


class SpatialLDOS:
    """Holds the results of :meth:`KPM.calc_spatial_ldos` (data: energy x site).

    Same surface as the reference class (pybinding/chebyshev.py:20-62): a product of a `Series` and a structure map."""

    def __init__(self, data, energy, structure):
        self.data = data
        self.energy = energy
        self.structure = structure  # the selected sites: a sliced pybinding System, or a `SiteSelection`

    def structure_map(self, energy):
        """The LDOS of every selected site at the sampled energy closest to `energy`, attached to the structure"""
        idx = np.argmin(abs(self.energy - energy))
        return self.structure.with_data(self.data[idx])

    def ldos(self, position, sublattice=""):
        """The LDOS as a function of energy at the selected site closest to `position` -> :class:`Series`"""
        idx = self.structure.find_nearest(position, sublattice)
        return results.Series(self.energy, self.data[:, idx],
                              labels=dict(variable="E (eV)", data="LDOS", columns="orbitals"))

    def ldos_at(self, energy):
        """LDOS of every selected site at the sampled energy closest to `energy` (plain array)"""
        idx = np.argmin(abs(self.energy - energy))
        return self.data[idx]


class KPMKernel:
    """Damping kernel (reference: kpm::Kernel bound as `KPMKernel`, cppmodule/src/kpm.cpp:68-74)"""

    def __init__(self, kind, lambda_value=4.0):
        if kind == _lib.LORENTZ and lambda_value <= 0:
            raise ValueError("Lorentz kernel: lambda must be positive.")
        self.kind = kind
        self.lambda_value = float(lambda_value)

    def damping_coefficients(self, num_moments):
        out = np.zeros(int(num_moments))
        _lib.load().pbk_kernel_damping(self.kind, self.lambda_value, int(num_moments), _lib.ptr(out))
        return out

    def required_num_moments(self, scaled_broadening):
        out = C.c_int32(0)
        _lib.load().pbk_kernel_required_num_moments(self.kind, self.lambda_value, float(scaled_broadening), C.byref(out))
        return out.value


def jackson_kernel():
    """The Jackson kernel -- a good general-purpose kernel, appropriate for most applications"""
    return KPMKernel(_lib.JACKSON)


def lorentz_kernel(lambda_value=4.0):
    """The Lorentz kernel -- best for Green's function"""
    return KPMKernel(_lib.LORENTZ, lambda_value)


def dirichlet_kernel():
    """The Dirichlet kernel -- returns raw moments, least favorable choice"""
    return KPMKernel(_lib.DIRICHLET)


class KPMStats:
    """Attribute view of `pbk_stats` (reference: `KPMStats`, cppmodule/src/kpm.cpp:50-66)"""

    def __init__(self, raw):
        for name, _ in raw._fields_:
            setattr(self, name, getattr(raw, name))
        self.uses_full_system = bool(self.uses_full_system)

    @property
    def ops(self):
        operations = self.nnz * 2 + self.vec * 5
        return self.multiplier * operations / self.moments_time if self.moments_time > 0 else 0.0

    def as_dict(self):
        return dict(self.__dict__)


class Deferred:
    """Lazily computed result (reference: `Deferred<T>`, cppmodule/include/thread.hpp:9-43)"""

    def __init__(self, solver, compute):
        self.solver = solver
        self._compute = compute
        self._result = None
        self._done = False
        self._lock = threading.Lock()

    def compute(self):
        with self._lock:
            if not self._done:
                self._result = self._compute()
                self._done = True

    @property
    def result(self):
        self.compute()
        return self._result


class _CudaImpl:
    """`cpb::KPM` facade over one `pbk_ctx` (one GPU).  Thread-safe per object like the reference."""

    def __init__(self, model, energy_range=(0, 0), kernel=None, matrix_format="ELL", optimal_size=True,
                 interleaved=True, lanczos_precision=0.002, num_threads=0, progress_callback=None,
                 device=0, max_batch=0, locality_tile=0):
        self._lib = _lib.load()
        kernel = kernel or jackson_kernel()
        emin, emax = (float(energy_range[0]), float(energy_range[1])) if energy_range else (0.0, 0.0)
        cfg = _lib.Config(np.float32(emin), np.float32(emax), kernel.kind, kernel.lambda_value,
                          int(bool(optimal_size)), int(bool(interleaved)), int(matrix_format == "ELL"),
                          np.float32(lanczos_precision), int(max_batch), int(locality_tile))
        self._handle = C.c_void_p()
        status = self._lib.pbk_create(C.byref(self._handle), int(device), C.byref(cfg))
        _lib.raise_for(status, None)
        self._kernel = kernel
        self._progress_ref = None
        if progress_callback is not None:
            def trampoline(delta, total, _user):
                progress_callback(delta, total)
            self._progress_ref = _lib.PROGRESS_FN(trampoline)
            self._lib.pbk_set_progress_callback(self._handle, self._progress_ref, None)
        self._model = None
        self._keep = None
        self.model = model

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle:
            self._lib.pbk_destroy(handle)
            self._handle = None

    def _check(self, status):
        _lib.raise_for(status, self._handle)

    # -- model / Hamiltonian ------------------------------------------------------------------------
    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, model):
        model = model.eval() if hasattr(model, "eval") and model.eval() is not None else model
        h = model.hamiltonian.tocsr()
        if not h.has_sorted_indices:
            h = h.copy()
            h.sort_indices()
        dtype = np.dtype(h.dtype)
        if dtype not in _lib.DTYPES:
            raise TypeError("unsupported Hamiltonian dtype {}".format(dtype))
        if h.nnz >= 2**31:
            raise ValueError("the Hamiltonian has too many non-zeros for int32 indices")
        indptr = np.ascontiguousarray(h.indptr, np.int32)
        indices = np.ascontiguousarray(h.indices, np.int32)
        data = np.ascontiguousarray(h.data)
        self._check(self._lib.pbk_set_hamiltonian(self._handle, _lib.DTYPES[dtype], h.shape[0],
                                                  _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data)))
        self._model = model
        self._size = h.shape[0]
        self._dtype = dtype

    @property
    def system(self):
        return self._model.system

    @property
    def scaling_factors(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.pbk_scaling_factors(self._handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def bounds(self):
        mn, mx, loops = C.c_double(), C.c_double(), C.c_int32()
        self._check(self._lib.pbk_bounds(self._handle, C.byref(mn), C.byref(mx), C.byref(loops)))
        return mn.value, mx.value, loops.value

    @property
    def kernel(self):
        return self._kernel

    @property
    def stats(self):
        raw = _lib.Stats()
        self._check(self._lib.pbk_get_stats(self._handle, C.byref(raw)))
        return KPMStats(raw)

    def report(self, shortform=False):
        buf = C.create_string_buffer(2048)
        self._check(self._lib.pbk_report(self._handle, int(shortform), buf, 2048))
        return buf.value.decode()

    # -- multi-GPU ----------------------------------------------------------------------------------
    def comm_init(self, world_size, rank, unique_id):
        self._check(self._lib.pbk_comm_init(self._handle, int(world_size), int(rank), unique_id))

    # -- raw moments (compute-strategy level) -----------------------------------------------------------
    def moments_dos(self, num_moments, num_random):
        out = np.zeros(num_moments, np.complex128)
        self._check(self._lib.pbk_moments_dos(self._handle, num_moments, num_random, _lib.ptr(out)))
        return out

    def moments_ldos(self, num_moments, indices):
        idx = np.ascontiguousarray(np.atleast_1d(indices), np.int32)
        out = np.zeros((num_moments, idx.size), np.complex128)
        self._check(self._lib.pbk_moments_ldos(self._handle, num_moments, _lib.ptr(idx), idx.size, _lib.ptr(out)))
        return out

    def moments_greens(self, num_moments, row, cols):
        cols = np.ascontiguousarray(np.atleast_1d(cols), np.int32)
        out = np.zeros((cols.size, num_moments), np.complex128)
        self._check(self._lib.pbk_moments_greens(self._handle, num_moments, int(row), _lib.ptr(cols), cols.size,
                                                 _lib.ptr(out)))
        return out

    def moments_kubo(self, num_moments, left, right, num_random):
        left = np.ascontiguousarray(left, np.float32)
        right = np.ascontiguousarray(right, np.float32)
        out = np.zeros((num_moments, num_moments), np.complex128)
        self._check(self._lib.pbk_moments_kubo(self._handle, num_moments, _lib.ptr(left), _lib.ptr(right),
                                               num_random, _lib.ptr(out)))
        return out

    def moments_diagonal(self, num_moments, vectors):
        v = np.ascontiguousarray(np.atleast_2d(vectors), np.complex128)
        out = np.zeros((num_moments, v.shape[0]), np.complex128)
        self._check(self._lib.pbk_moments_diagonal(self._handle, num_moments, _lib.ptr(v), v.shape[0], _lib.ptr(out)))
        return out

    def random_vectors(self, count):
        out = np.zeros((count, self._size), np.complex128)
        self._check(self._lib.pbk_random_vectors(self._handle, count, _lib.ptr(out)))
        return out if self._dtype.kind == "c" else out.real.copy()

    # -- cpb::KPM facade (cppcore/src/KPM.cpp) ------------------------------------------------------------
    def moments(self, num_moments, alpha, beta, op):
        size = self._size
        alpha = np.atleast_1d(np.asarray(alpha))
        beta = np.atleast_1d(np.asarray(beta)) if beta is not None else np.zeros(0)
        op_size = 0 if op is None else int(np.prod(op.shape))
        op_ok = op is None or op_size == 0 or op.shape == (size, size)
        for name, ok in (("alpha", alpha.size == size), ("beta", beta.size in (0, size)), ("operator", op_ok)):
            if not ok:
                raise RuntimeError("Size mismatch between the model Hamiltonian and the given "
                                   "argument '{}'".format(name))
        has_op = op is not None and op_size > 1
        if self._dtype.kind != "c":
            checks = (("alpha", np.iscomplexobj(alpha) and np.any(alpha.imag != 0)),
                      ("beta", np.iscomplexobj(beta) and np.any(beta.imag != 0)),
                      ("operator", has_op and np.iscomplexobj(op.data) and np.any(op.data.imag != 0)))
            for name, bad in checks:
                if bad:
                    raise RuntimeError("The model Hamiltonian is real, but the given argument "
                                       "'{}' is complex".format(name))
        a = np.ascontiguousarray(alpha, np.complex128)
        b = np.ascontiguousarray(beta, np.complex128) if beta.size else None
        op_rows, ip, ix, dt = 0, None, None, None
        if has_op:
            op = op.tocsr()
            op.sort_indices()
            op_rows = op.shape[0]
            ip = np.ascontiguousarray(op.indptr, np.int32)
            ix = np.ascontiguousarray(op.indices, np.int32)
            dt = np.ascontiguousarray(op.data, np.complex128)
        out = np.zeros(int(num_moments), np.complex128)
        self._check(self._lib.pbk_moments(self._handle, int(num_moments), _lib.ptr(a), _lib.ptr(b), op_rows,
                                          _lib.ptr(ip), _lib.ptr(ix), _lib.ptr(dt), _lib.ptr(out)))
        return out

    def calc_greens(self, i, j, energy, broadening):
        size = self._size
        single = np.isscalar(j)
        cols = np.ascontiguousarray(np.atleast_1d(j), np.int32)
        if i < 0 or i >= size or np.any(cols < 0) or np.any(cols >= size):
            raise RuntimeError("KPM::calc_greens(i,j): invalid value for i or j.")
        e = np.ascontiguousarray(energy, np.float64)
        out = np.zeros((cols.size, e.size), np.complex128)
        self._check(self._lib.pbk_calc_greens(self._handle, int(i), _lib.ptr(cols), cols.size, _lib.ptr(e), e.size,
                                              float(broadening), _lib.ptr(out)))
        return out[0] if single else [row.copy() for row in out]

    def _ldos_indices(self, idx, energy, broadening):
        e = np.ascontiguousarray(energy, np.float64)
        idx = np.ascontiguousarray(np.atleast_1d(idx), np.int32)
        out = np.zeros((idx.size, e.size))
        self._check(self._lib.pbk_calc_ldos(self._handle, _lib.ptr(e), e.size, float(broadening), _lib.ptr(idx),
                                            idx.size, _lib.ptr(out)))
        return np.asfortranarray(out.T)  # energy x index, column-major like ArrayXXdCM

    def calc_ldos(self, energy, broadening, position, sublattice="", reduce=True):
        system_index = self._model.system.find_nearest(position, sublattice)
        ham_idx = self._model.system.to_hamiltonian_indices(system_index)
        result = self._ldos_indices(ham_idx, energy, broadening)
        return result.sum(axis=1, keepdims=True) if (reduce and result.shape[1] > 1) else result

    def calc_spatial_ldos(self, energy, broadening, shape, sublattice=""):
        if getattr(self._model, "is_multiorbital", False):
            raise RuntimeError("This function doesn't currently support multi-orbital models")
        system = self._model.system
        contains = np.asarray(shape.contains(*system.positions))
        start, end = sublattice_range(system, sublattice)
        indices = start + np.flatnonzero(contains[start:end])
        return self._ldos_indices(indices, energy, broadening)

    def calc_dos(self, energy, broadening, num_random):
        e = np.ascontiguousarray(energy, np.float64)
        out = np.zeros(e.size)
        self._check(self._lib.pbk_calc_dos(self._handle, _lib.ptr(e), e.size, float(broadening), int(num_random),
                                           _lib.ptr(out)))
        return out

    def calc_conductivity(self, chemical_potential, broadening, temperature, direction, num_random, num_points):
        if len(direction) != 2 or any(d not in "xyz" for d in direction):
            raise RuntimeError("Invalid direction: must be 'xx', 'xy', 'zz', or similar.")
        system = self._model.system
        p = system.expanded_positions if getattr(self._model, "is_multiorbital", False) else system.positions
        axes = dict(x=p.x, y=p.y, z=p.z)
        left = np.ascontiguousarray(axes[direction[0]], np.float32)
        right = np.ascontiguousarray(axes[direction[1]], np.float32)
        mu = np.ascontiguousarray(chemical_potential, np.float64)
        out = np.zeros(mu.size, np.complex128)
        self._check(self._lib.pbk_calc_conductivity(self._handle, _lib.ptr(left), _lib.ptr(right), _lib.ptr(mu),
                                                    mu.size, float(broadening), float(temperature),
                                                    int(num_random), int(num_points), _lib.ptr(out)))
        return out.real.copy()

    def deferred_ldos(self, energy, broadening, position, sublattice=""):
        energy = np.array(energy, np.float64)
        return Deferred(self, lambda: self.calc_ldos(energy, broadening, position, sublattice))


class KPM:
    """The common interface for various KPM implementations

    It should not be created directly but via specific functions like :func:`kpm`.
    Same methods, arguments and return types as `pybinding.chebyshev.KPM` (chebyshev.py:65-317).
    """

    def __init__(self, impl):
        if hasattr(impl, "hamiltonian") and hasattr(impl, "system"):
            raise TypeError("You're probably looking for `pb.kpm()` (lowercase).")
        self.impl = impl

    @property
    def model(self):
        """The tight-binding model holding the Hamiltonian"""
        return self.impl.model

    @model.setter
    def model(self, model):
        self.impl.model = model

    @property
    def system(self):
        """The tight-binding system (shortcut for `KPM.model.system`)"""
        return self.impl.system

    @property
    def scaling_factors(self) -> tuple:
        """A tuple of KPM scaling factors `a` and `b`"""
        return self.impl.scaling_factors

    @property
    def kernel(self):
        """The damping kernel"""
        return self.impl.kernel

    @property
    def stats(self):
        return self.impl.stats

    def report(self, shortform=False):
        """Return a report of the last computation"""
        return self.impl.report(shortform)

    def __call__(self, *args, **kwargs):
        warnings.warn("Use .calc_greens() instead", DeprecationWarning)
        return self.calc_greens(*args, **kwargs)

    def moments(self, num_moments, alpha, beta=None, op=None):
        r"""Calculate KPM moments in the form of expectation values :math:`\mu_n = <\beta|op \cdot T_n(H)|\alpha>`

        Returned moments are damped by the kernel and `mu_0` carries the 1/2 factor, like the reference.
        """
        if beta is None:
            beta = []
        if op is not None:
            op = op.tocsr()
        return self.impl.moments(num_moments, alpha, beta, op)

    def calc_greens(self, i, j, energy, broadening):
        """Calculate Green's function of a single Hamiltonian element (or a list of `j` elements)"""
        return self.impl.calc_greens(i, j, energy, broadening)

    def calc_ldos(self, energy, broadening, position, sublattice="", reduce=True):
        """Calculate the local density of states as a function of energy -> :class:`Series`"""
        ldos = self.impl.calc_ldos(energy, broadening, position, sublattice, reduce)
        return results.Series(energy, ldos.squeeze(), labels=dict(variable="E (eV)", data="LDOS",
                                                                  columns="orbitals"))

    def calc_spatial_ldos(self, energy, broadening, shape, sublattice=""):
        """Calculate the LDOS as a function of energy and space (in the area of the given shape)"""
        ldos = self.impl.calc_spatial_ldos(energy, broadening, shape, sublattice)
        system = self.system
        contains = np.asarray(shape.contains(*system.positions))
        if hasattr(system, "__getitem__") and hasattr(system, "sub"):   # a pybinding System: slice it like the reference does
            smap = system[contains]
            if sublattice:
                smap = smap[smap.sub == sublattice]
            return SpatialLDOS(ldos, np.asarray(energy), smap)
        start, end = sublattice_range(system, sublattice)
        return SpatialLDOS(ldos, np.asarray(energy), SiteSelection(system, start + np.flatnonzero(contains[start:end])))

    def calc_dos(self, energy, broadening, num_random=1):
        """Calculate the density of states as a function of energy -> :class:`Series`"""
        dos = self.impl.calc_dos(energy, broadening, num_random)
        return results.Series(energy, dos, labels=dict(variable="E (eV)", data="DOS"))

    def deferred_ldos(self, energy, broadening, position, sublattice=""):
        """Same as :meth:`calc_ldos` but for parallel computation -> :class:`Deferred`"""
        return self.impl.deferred_ldos(energy, broadening, position, sublattice)

    def calc_conductivity(self, chemical_potential, broadening, temperature,
                          direction="xx", volume=1.0, num_random=1, num_points=1000):
        """Calculate Kubo-Bastin electrical conductivity as a function of chemical potential"""
        data = self.impl.calc_conductivity(chemical_potential, broadening, temperature,
                                           direction, num_random, num_points)
        if volume != 1.0:
            data /= volume
        return results.Series(chemical_potential, data,
                              labels=dict(variable=r"$\mu$ (eV)", data=r"$\sigma (e^2/h)$"))


class _ComputeProgressReporter:
    def __call__(self, delta, total):
        if total == 1:
            return  # Skip reporting for short jobs
        if delta < 0:
            print("Computing KPM moments...")
            self.done = 0
        elif delta == total:
            print("\rKPM moments: 100%")
        else:
            self.done = getattr(self, "done", 0) + delta
            print("\rKPM moments: {:.0f}%".format(100 * self.done / total), end="", flush=True)


def kpm(model, energy_range=None, kernel="default", num_threads="auto", silent=False, **kwargs):
    """The B200 implementation of the Kernel Polynomial Method (signature of `pybinding.kpm`)

    Parameters
    ----------
    model : Model
        Anything with `.hamiltonian` (scipy CSR) and `.system` (see `pybinding_b200.synthetic`), e.g. a `pb.Model`.
    energy_range : Optional[Tuple[float, float]]
        `(min, max)` eigenvalue bounds; found with the Lanczos procedure when omitted.
    kernel : Kernel
        :func:`jackson_kernel` (default), :func:`lorentz_kernel` or :func:`dirichlet_kernel`.
    num_threads : int
        Accepted for compatibility; the GPU engine has no CPU worker threads.
    silent : bool
        Don't show any progress messages.
    **kwargs
        `matrix_format`, `optimal_size`, `interleaved`, `lanczos_precision`, `progress_callback` as in the
        reference (cppmodule/src/kpm.cpp:27-36), plus `device` (CUDA ordinal), `max_batch` and
        `binding`: "ctypes" (default; `_CudaImpl` over the C ABI) or "pybind11" (the compiled `_pbkpm` module, the
        C++ counterpart of `_pybinding.kpm`, csrc/pymodule.cpp).  Both drive the same `libpbkpm.so`.
    """
    binding = kwargs.pop("binding", "ctypes")
    if kernel != "default":
        kwargs["kernel"] = kernel
    if num_threads != "auto":
        kwargs["num_threads"] = num_threads
    if "progress_callback" not in kwargs:
        kwargs["progress_callback"] = _ComputeProgressReporter()
    if silent:
        del kwargs["progress_callback"]
    if binding == "pybind11":
        from . import _pbkpm   # ImportError if the extension was not built: there is no fallback
        return KPM(_pbkpm.kpm(model, tuple(energy_range) if energy_range is not None else (0, 0), **kwargs))
    if binding != "ctypes":
        raise ValueError("binding must be 'ctypes' or 'pybind11'")
    return KPM(_CudaImpl(model, energy_range or (0, 0), **kwargs))


def kpm_cuda(model, energy_range=None, kernel="default", **kwargs):
    """Same as :func:`kpm` (the reference's name for a GPU implementation, chebyshev.py:381-404)"""
    kwargs.setdefault("silent", True)
    return kpm(model, energy_range, kernel, **kwargs)
