"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product (`pybinding_b200`) never does.

`OracleKPM` mirrors `kpm::Core` of the reference (cppcore/src/kpm/Core.cpp:35-156): the same
quantities, computed by the C++ restatement in `kpm_oracle.cpp`.  `hp=True` selects double-precision
accumulation + reconstruction (the arithmetic the GPU engine implements); `hp=False` is the
reference-faithful native-precision mode.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkpm_oracle.so")
_lib = None

DTYPES = {np.dtype(np.float32): 0, np.dtype(np.complex64): 1,
          np.dtype(np.float64): 2, np.dtype(np.complex128): 3}
KERNELS = {"jackson": 0, "lorentz": 1, "dirichlet": 2}


def build(force=False):
    src = os.path.join(_HERE, "kpm_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_error.restype = C.c_char_p
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                    C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_moments_seconds.restype = C.c_double
        _lib.orc_moments_seconds.argtypes = [C.c_void_p]
        for name in ("orc_bounds", "orc_required_num_moments", "orc_optimize_for", "orc_oh_get", "orc_oh_matrix",
                     "orc_oh_slice_index", "orc_random_vectors", "orc_dos_moments", "orc_ldos_moments",
                     "orc_greens_moments", "orc_kubo_moments", "orc_moments", "orc_calc_dos", "orc_calc_ldos",
                     "orc_calc_greens", "orc_calc_conductivity", "orc_last_num_moments", "orc_time_dos",
                     "orc_time_dos_probe"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(status):
    if status != 0:
        raise RuntimeError(lib().orc_error().decode())


def scale(emin, emax):
    """(a, b) of kpm::Scale (Bounds.hpp:11-33), float quirks included"""
    out = np.zeros(2)
    lib().orc_scale(C.c_double(emin), C.c_double(emax), _p(out))
    return float(out[0]), float(out[1])


def damping_coefficients(kernel, n, lambda_value=4.0):
    out = np.zeros(n)
    lib().orc_kernel_damping(KERNELS[kernel], C.c_double(lambda_value), C.c_int(n), _p(out))
    return out


def required_num_moments(kernel, scaled_broadening, lambda_value=4.0):
    return lib().orc_kernel_required_num_moments(KERNELS[kernel], C.c_double(lambda_value),
                                                 C.c_double(scaled_broadening))


def hardware_threads():
    return lib().orc_hardware_threads()


class OracleKPM:
    def __init__(self, hamiltonian, energy_range=(0, 0), kernel="jackson", lambda_value=4.0,
                 matrix_format="ELL", optimal_size=True, interleaved=True, lanczos_precision=0.002,
                 num_threads=1, hp=False):
        h = hamiltonian.tocsr()
        h.sort_indices()
        self.dtype = np.dtype(h.dtype)
        self.n = h.shape[0]
        self._keep = (np.ascontiguousarray(h.indptr, np.int32), np.ascontiguousarray(h.indices, np.int32),
                      np.ascontiguousarray(h.data))
        energy_range = energy_range or (0, 0)
        self.handle = lib().orc_create(DTYPES[self.dtype], self.n, _p(self._keep[0]), _p(self._keep[1]),
                                       _p(self._keep[2]), float(energy_range[0]), float(energy_range[1]),
                                       KERNELS[kernel], float(lambda_value), int(matrix_format == "ELL"),
                                       int(optimal_size), int(interleaved), float(lanczos_precision),
                                       int(num_threads), int(hp))
        if not self.handle:
            raise ValueError(lib().orc_error().decode())
        self.handle = C.c_void_p(self.handle)

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_destroy(self.handle)
            self.handle = None

    # -- bounds / scale -------------------------------------------------------------------------
    def bounds(self):
        out = np.zeros(5)
        _check(lib().orc_bounds(self.handle, _p(out)))
        return dict(min=out[0], max=out[1], a=out[2], b=out[3], loops=int(out[4]))

    @property
    def scaling_factors(self):
        b = self.bounds()
        return b["a"], b["b"]

    def required_num_moments(self, broadening):
        m = lib().orc_required_num_moments(self.handle, C.c_double(broadening))
        if m < 0:
            _check(1)
        return m

    # -- optimized Hamiltonian test hooks -----------------------------------------------------------
    def optimize_for(self, src, dest):
        src = np.ascontiguousarray(np.atleast_1d(src), np.int32)
        dest = np.ascontiguousarray(np.atleast_1d(dest), np.int32)
        info = np.zeros(4, np.int32)
        _check(lib().orc_optimize_for(self.handle, _p(src), src.size, _p(dest), dest.size, _p(info)))
        nslices, src_off, dest_off, nnz = info.tolist()
        slices = np.zeros(nslices, np.int32)
        osrc = np.zeros(src.size, np.int32)
        odest = np.zeros(dest.size, np.int32)
        rmap = np.zeros(self.n, np.int32)
        _check(lib().orc_oh_get(self.handle, _p(slices), _p(osrc), _p(odest), _p(rmap)))
        return dict(slices=slices, src=osrc, dest=odest, src_offset=src_off, dest_offset=dest_off,
                    nnz=nnz, reorder_map=rmap)

    def optimized_matrix(self, nnz):
        import scipy.sparse as sp
        indptr = np.zeros(self.n + 1, np.int32)
        indices = np.zeros(nnz, np.int32)
        data = np.zeros(nnz, np.complex128)
        _check(lib().orc_oh_matrix(self.handle, _p(indptr), _p(indices), _p(data)))
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.n))

    def slice_index(self, n, num_moments):
        return lib().orc_oh_slice_index(self.handle, n, num_moments)

    # -- raw moments ----------------------------------------------------------------------------------
    def random_vectors(self, count):
        """The first `count` starter vectors of the reference's random stream, original site order"""
        out = np.zeros((count, self.n), np.complex128)
        _check(lib().orc_random_vectors(self.handle, count, _p(out)))
        return out if self.dtype.kind == "c" else out.real.copy()

    def dos_moments(self, num_moments, num_random):
        out = np.zeros(num_moments, np.complex128)
        _check(lib().orc_dos_moments(self.handle, num_moments, num_random, _p(out)))
        return out

    def ldos_moments(self, num_moments, idx):
        idx = np.ascontiguousarray(np.atleast_1d(idx), np.int32)
        out = np.zeros((num_moments, idx.size), np.complex128)
        _check(lib().orc_ldos_moments(self.handle, num_moments, _p(idx), idx.size, _p(out)))
        return out

    def greens_moments(self, num_moments, row, cols):
        cols = np.ascontiguousarray(np.atleast_1d(cols), np.int32)
        out = np.zeros((cols.size, num_moments), np.complex128)
        _check(lib().orc_greens_moments(self.handle, num_moments, int(row), _p(cols), cols.size, _p(out)))
        return out

    def kubo_moments(self, num_moments, left, right, num_random):
        left = np.ascontiguousarray(left, np.float32)
        right = np.ascontiguousarray(right, np.float32)
        out = np.zeros((num_moments, num_moments), np.complex128)
        _check(lib().orc_kubo_moments(self.handle, num_moments, _p(left), _p(right), num_random, _p(out)))
        return out

    def moments(self, num_moments, alpha, beta=None, op=None):
        alpha = np.ascontiguousarray(alpha, np.complex128)
        beta_p = None
        if beta is not None and len(beta) != 0:
            beta = np.ascontiguousarray(beta, np.complex128)
            beta_p = _p(beta)
        op_rows, ip, ix, dt = 0, None, None, None
        if op is not None and op.shape[0] > 1:
            op = op.tocsr()
            op.sort_indices()
            op_rows = op.shape[0]
            k_ip = np.ascontiguousarray(op.indptr, np.int32)
            k_ix = np.ascontiguousarray(op.indices, np.int32)
            k_dt = np.ascontiguousarray(op.data, np.complex128)
            ip, ix, dt = _p(k_ip), _p(k_ix), _p(k_dt)
        out = np.zeros(num_moments, np.complex128)
        _check(lib().orc_moments(self.handle, num_moments, _p(alpha), beta_p, op_rows, ip, ix, dt, _p(out)))
        return out

    # -- full calculations ------------------------------------------------------------------------------
    def calc_dos(self, energy, broadening, num_random=1):
        e = np.ascontiguousarray(energy, np.float64)
        out = np.zeros(e.size)
        _check(lib().orc_calc_dos(self.handle, _p(e), e.size, C.c_double(broadening), num_random, _p(out)))
        return out

    def calc_ldos(self, energy, broadening, idx):
        e = np.ascontiguousarray(energy, np.float64)
        idx = np.ascontiguousarray(np.atleast_1d(idx), np.int32)
        out = np.zeros((idx.size, e.size))
        _check(lib().orc_calc_ldos(self.handle, _p(e), e.size, C.c_double(broadening), _p(idx), idx.size, _p(out)))
        return out.T  # energy x index

    def calc_greens(self, row, cols, energy, broadening):
        e = np.ascontiguousarray(energy, np.float64)
        single = np.isscalar(cols)
        c = np.ascontiguousarray(np.atleast_1d(cols), np.int32)
        out = np.zeros((c.size, e.size), np.complex128)
        _check(lib().orc_calc_greens(self.handle, int(row), _p(c), c.size, _p(e), e.size, C.c_double(broadening), _p(out)))
        return out[0] if single else list(out)

    def calc_conductivity(self, chemical_potential, broadening, temperature, left, right, num_random=1,
                          num_points=1000):
        mu = np.ascontiguousarray(chemical_potential, np.float64)
        left = np.ascontiguousarray(left, np.float32)
        right = np.ascontiguousarray(right, np.float32)
        out = np.zeros(mu.size, np.complex128)
        _check(lib().orc_calc_conductivity(self.handle, _p(left), _p(right), _p(mu), mu.size, C.c_double(broadening),
                                           C.c_double(temperature), num_random, num_points, _p(out)))
        return out.real

    @property
    def last_num_moments(self):
        return lib().orc_last_num_moments(self.handle)

    @property
    def moments_seconds(self):
        return lib().orc_moments_seconds(self.handle)

    def time_dos(self, num_moments, num_random, num_threads, cheap_starter=False):
        """Seconds spent in the reference-shaped DOS moment computation (cpu_baseline).

        cheap_starter=True replaces the reference's serial MT19937 (+ complex exp) starter by a trivial +-1 fill,
        so that a short sample measures the recursion throughput instead of the fixed per-vector starter cost."""
        t = C.c_double(0)
        _check(lib().orc_time_dos(self.handle, num_moments, num_random, num_threads, int(cheap_starter), C.byref(t)))
        return t.value

    def time_dos_probe(self, num_moments, num_random, num_threads, n1, n2, cheap_starter=False):
        """One reference-shaped DOS moment run with a probe around recursion steps n1 .. n2 (2 moments per step).

        Returns (total_seconds, probe_seconds, jobs): total = the whole moments phase including the per-vector fixed cost
        (starters, allocation / first touch of the vector blocks, r1); probe = wall time of the slowest thread-pool job
        between the two steps, i.e. the asymptotic cost of 2 * (n2 - n1) moments for all `num_random` vectors when
        the jobs run as one wave (num_random <= num_threads * SIMD batch)."""
        out = np.zeros(3)
        _check(lib().orc_time_dos_probe(self.handle, num_moments, num_random, num_threads, int(cheap_starter),
                                        int(n1), int(n2), _p(out)))
        return float(out[0]), float(out[1]), int(out[2])
