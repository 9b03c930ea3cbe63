"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol of include/pbkpm.h."""
import os
import re

import numpy as np
import pytest

import pybinding_b200 as pb
from pybinding_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pbkpm.h")).read()
    declared = sorted(set(re.findall(r"\b(pbk_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pbk_version() == 100


def test_kernels_match_reference_formulas():
    """Kernel.cpp:6-49 (float pi), Kernel.hpp:9-13 (4k+2 rounding)"""
    pi_f = float(np.float32(np.pi))
    g = pb.jackson_kernel().damping_coefficients(10)
    n = np.arange(10.0)
    expected = ((11 - n) * np.cos(pi_f * n / 11) + np.sin(pi_f * n / 11) / np.tan(pi_f / 11)) / 11
    assert np.allclose(g, expected, rtol=1e-14)
    assert pb.jackson_kernel().required_num_moments(pi_f / 1024) == 1026
    assert pb.lorentz_kernel(4.0).required_num_moments(0.15 / 8.5) == 230
    assert np.all(pb.dirichlet_kernel().damping_coefficients(5) == 1)
    lam = 4.0
    g = pb.lorentz_kernel(lam).damping_coefficients(8)
    assert np.allclose(g, np.sinh(lam * (1 - np.arange(8.0) / 8)) / np.sinh(lam), rtol=1e-14)
    with pytest.raises(ValueError):
        pb.lorentz_kernel(-1)


def test_no_cpu_fallback_without_device():
    """Creating a context without a usable GPU must fail loudly (skipped where a GPU exists)"""
    import ctypes as C
    count = C.c_int(0)
    _lib.load().pbk_device_count(C.byref(count))
    if count.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.PbkError) as excinfo:
        pb.kpm(pb.graphene_rectangle(2), silent=True)
    assert "no CPU fallback" in str(excinfo.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pybinding_b200/ may reference it"""
    pkg = os.path.join(ROOT, "pybinding_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "_never_", os.path.join(dirpath, f)


def test_pybind11_module_mirrors_the_reference_binding():
    """`_pbkpm` (csrc/pymodule.cpp) is the compiled counterpart of `_pybinding.kpm` / `KPM` (cppmodule/src/kpm.cpp:8-102)"""
    from pybinding_b200 import _pbkpm
    assert _pbkpm.abi_version == _lib.load().pbk_version()
    for name in ("kpm", "kpm_cuda", "KPM", "KPMKernel", "KPMStats", "DeferredXd", "jackson_kernel", "lorentz_kernel", "dirichlet_kernel"):
        assert hasattr(_pbkpm, name), name
    for name in ("moments", "calc_greens", "calc_dos", "calc_conductivity", "calc_ldos", "calc_spatial_ldos", "deferred_ldos",
                 "report", "model", "system", "scaling_factors", "kernel", "stats"):     # cppmodule/src/kpm.cpp:76-102
        assert hasattr(_pbkpm.KPM, name), name
    # kernels are context-free: same numbers as the ctypes layer
    assert np.array_equal(_pbkpm.jackson_kernel().damping_coefficients(12), pb.jackson_kernel().damping_coefficients(12))
    assert _pbkpm.lorentz_kernel(3.0).required_num_moments(0.01) == pb.lorentz_kernel(3.0).required_num_moments(0.01)
    with pytest.raises(ValueError):
        _pbkpm.lorentz_kernel(0.0)
    import ctypes as C
    count = C.c_int(0)
    _lib.load().pbk_device_count(C.byref(count))
    if count.value == 0:   # std::runtime_error -> RuntimeError, like the reference's exceptions through pybind11
        with pytest.raises(RuntimeError) as excinfo:
            pb.kpm(pb.graphene_rectangle(2), silent=True, binding="pybind11")
        assert "no CPU fallback" in str(excinfo.value)
    with pytest.raises(ValueError):
        pb.kpm(pb.graphene_rectangle(2), silent=True, binding="nope")
