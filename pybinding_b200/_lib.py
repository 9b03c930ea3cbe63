"""ctypes binding of libpbkpm.so (the C ABI in include/pbkpm.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, creating a
KPM object raises.  `load()` itself only needs the library file, so symbol checks work without a GPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpbkpm.so")

F32, C64, F64, C128 = 0, 1, 2, 3
DTYPES = {np.dtype(np.float32): F32, np.dtype(np.complex64): C64,
          np.dtype(np.float64): F64, np.dtype(np.complex128): C128}
OK, INVALID_ARGUMENT, RUNTIME_ERROR, LOGIC_ERROR, CUDA_ERROR, NCCL_ERROR = range(6)
JACKSON, LORENTZ, DIRICHLET = 0, 1, 2


class PbkError(RuntimeError):
    """CUDA / NCCL failures of the engine"""


class Config(C.Structure):
    _fields_ = [("min_energy", C.c_float), ("max_energy", C.c_float), ("kernel", C.c_int32),
                ("lambda_value", C.c_double), ("optimal_size", C.c_int32), ("interleaved", C.c_int32),
                ("matrix_format", C.c_int32), ("lanczos_precision", C.c_float), ("max_batch", C.c_int32),
                ("locality_tile", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("num_moments", C.c_int64), ("uses_full_system", C.c_int32), ("nnz", C.c_uint64),
                ("opt_nnz", C.c_uint64), ("vec", C.c_uint64), ("opt_vec", C.c_uint64),
                ("multiplier", C.c_double), ("matrix_memory", C.c_uint64), ("vector_memory", C.c_uint64),
                ("hamiltonian_time", C.c_double), ("moments_time", C.c_double), ("eps", C.c_double),
                ("kernel_launches", C.c_int64), ("step_launches", C.c_int64), ("step_ms", C.c_double),
                ("step_bytes", C.c_double), ("starter_ms", C.c_double), ("gemm_ms", C.c_double),
                ("gemm_flops", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("batch", C.c_int32), ("num_batches", C.c_int32), ("moments_device_ms", C.c_double),
                ("bulk_launches", C.c_int64),
                ("res_launches", C.c_int64), ("persist_launches", C.c_int64),
                ("graph_launches", C.c_int64)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_int64, C.c_int64, C.c_void_p)

#: every symbol declared in include/pbkpm.h (tests check that the library exports all of them)
SYMBOLS = [
    "pbk_version", "pbk_create", "pbk_destroy", "pbk_last_error", "pbk_set_progress_callback",
    "pbk_device_count", "pbk_set_hamiltonian", "pbk_bounds", "pbk_scaling_factors",
    "pbk_required_num_moments", "pbk_kernel_damping", "pbk_kernel_required_num_moments",
    "pbk_moments_dos", "pbk_moments_ldos", "pbk_moments_greens", "pbk_moments_kubo",
    "pbk_moments_diagonal", "pbk_random_vectors", "pbk_moments", "pbk_calc_dos", "pbk_calc_ldos",
    "pbk_calc_greens", "pbk_calc_conductivity", "pbk_get_stats", "pbk_report",
    "pbk_comm_unique_id", "pbk_comm_init", "pbk_comm_destroy", "pbk_locality_order", "pbk_locality_order2", "pbk_light_cone", "pbk_host_ell", "pbk_mt_jump_window", "pbk_shard",
]

_lib = None


def load():
    """Load libpbkpm.so; raises if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libpbkpm.so is missing at {}: build it with `make -C pybinding_b200/csrc` "
                          "(there is no CPU fallback)".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    lib.pbk_last_error.restype = C.c_char_p
    lib.pbk_last_error.argtypes = [C.c_void_p]
    lib.pbk_destroy.restype = None
    lib.pbk_destroy.argtypes = [C.c_void_p]
    lib.pbk_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Config)]
    lib.pbk_set_progress_callback.argtypes = [C.c_void_p, PROGRESS_FN, C.c_void_p]
    lib.pbk_set_hamiltonian.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pbk_bounds.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    lib.pbk_scaling_factors.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pbk_required_num_moments.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_int32)]
    lib.pbk_kernel_damping.argtypes = [C.c_int, C.c_double, C.c_int32, C.c_void_p]
    lib.pbk_kernel_required_num_moments.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int32)]
    lib.pbk_moments_dos.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.pbk_moments_ldos.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_moments_greens.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_moments_kubo.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_moments_diagonal.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_random_vectors.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_moments.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pbk_calc_dos.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_void_p]
    lib.pbk_calc_ldos.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_calc_greens.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_double, C.c_void_p]
    lib.pbk_calc_conductivity.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                          C.c_double, C.c_int32, C.c_int32, C.c_void_p]
    lib.pbk_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.pbk_report.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int64]
    lib.pbk_comm_unique_id.argtypes = [C.c_char_p]
    lib.pbk_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]
    lib.pbk_comm_destroy.argtypes = [C.c_void_p]
    lib.pbk_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.pbk_shard.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.pbk_mt_jump_window.argtypes = [C.c_uint64, C.c_void_p]
    lib.pbk_locality_order.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.pbk_host_ell.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pbk_light_cone.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    lib.pbk_locality_order2.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    _lib = lib
    return lib


def ptr(array):
    return None if array is None else array.ctypes.data_as(C.c_void_p)


def raise_for(status, handle):
    """Map a pbk_status to the exception type pybind11 gives the reference's C++ exceptions"""
    if status == OK:
        return
    msg = load().pbk_last_error(handle)
    msg = msg.decode() if msg else "pbkpm error {}".format(status)
    if status == INVALID_ARGUMENT:
        raise ValueError(msg)
    if status in (CUDA_ERROR, NCCL_ERROR):
        raise PbkError(msg)
    raise RuntimeError(msg)
