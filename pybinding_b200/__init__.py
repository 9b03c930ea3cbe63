"""pybinding_b200 -- pybinding's kernel-polynomial-method engine rebuilt for NVIDIA B200 (sm_100a).

Only the KPM hot path is provided: `kpm()` / `KPM` / kernels with the API of `pybinding.chebyshev`,
backed by libpbkpm.so (hand-written CUDA kernels behind the C ABI of include/pbkpm.h).
"""
from .chebyshev import (KPM, kpm, kpm_cuda, SpatialLDOS, Deferred, jackson_kernel, lorentz_kernel,
                        dirichlet_kernel)
from .results import Series
from . import synthetic
from . import parallel
from .synthetic import graphene_rectangle, cubic_anderson, Rectangle

__all__ = ["KPM", "kpm", "kpm_cuda", "SpatialLDOS", "Deferred", "jackson_kernel", "lorentz_kernel",
           "dirichlet_kernel", "Series", "synthetic", "parallel", "graphene_rectangle", "cubic_anderson", "Rectangle"]
