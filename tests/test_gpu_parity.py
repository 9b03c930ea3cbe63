"""GPU parity tests: the CUDA engine (through the C ABI / public API) vs the CPU oracle on identical inputs.

Tolerances (north_star): moments and curves within 1e-5 (f32/c64) and 1e-11 (f64/c128), measured relative to
the largest magnitude of the compared array.  The oracle runs in `hp` mode (vectors in the Hamiltonian's scalar
type, double-precision dot products / reconstruction) which is the arithmetic the GPU implements; the
reference-faithful native-f32 oracle's own distance to the same hp result is asserted to be no smaller than the
GPU's where that is informative (hard part 3).
"""
import numpy as np
import pytest

import pybinding_b200 as pb
from oracle.oracle import OracleKPM

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.complex64, np.float64, np.complex128]
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5,
       np.dtype(np.float64): 1e-11, np.dtype(np.complex128): 1e-11}


def rel_err(actual, expected):
    actual, expected = np.asarray(actual), np.asarray(expected)
    return float(np.abs(actual - expected).max() / np.abs(expected).max())


def make_model(dtype, width=8.0, onsite=0.3):
    dtype = np.dtype(dtype)
    return pb.graphene_rectangle(width, onsite=onsite, dtype=dtype,
                                 magnetic_field=400.0 if dtype.kind == "c" else 0.0)


@pytest.fixture(scope="module", params=DTYPES, ids=lambda d: np.dtype(d).name)
def setup(request):
    dtype = np.dtype(request.param)
    model = make_model(dtype)
    kpm = pb.kpm(model, energy_range=(-9.1, 9.3), silent=True)   # b != 0: exercises the diagonal insertion
    ref = OracleKPM(model.hamiltonian, energy_range=(-9.1, 9.3), hp=True)
    return dtype, model, kpm, ref


def test_random_starters_match_reference_stream(setup):
    dtype, model, kpm, ref = setup
    got = kpm.impl.random_vectors(3)
    expected = ref.random_vectors(3)
    if dtype.kind == "c":
        assert np.abs(got - expected).max() < (2e-7 if dtype == np.complex64 else 1e-15)
    else:
        assert np.array_equal(got, expected)  # +-1: bit exact


def test_scaling_factors(setup):
    dtype, model, kpm, ref = setup
    assert kpm.scaling_factors == pytest.approx(ref.scaling_factors, rel=0, abs=0)


@pytest.mark.parametrize("num_random", [1, 3, 9])
def test_dos_moments(setup, num_random):
    dtype, model, kpm, ref = setup
    M = 130
    got = kpm.impl.moments_dos(M, num_random)
    expected = ref.dos_moments(M, num_random)
    assert rel_err(got, expected) < TOL[dtype]


def test_dos_moments_explicit_starters(setup):
    """Identical caller-supplied starting vectors, all lanes advanced together"""
    dtype, model, kpm, ref = setup
    rng = np.random.default_rng(5)
    n = model.hamiltonian.shape[0]
    vecs = rng.standard_normal((5, n)) + (1j * rng.standard_normal((5, n)) if dtype.kind == "c" else 0)
    vecs = vecs.astype(dtype)
    M = 66
    got = kpm.impl.moments_diagonal(M, vecs)
    g = pb.dirichlet_kernel()
    ref_d = OracleKPM(model.hamiltonian, energy_range=(-9.1, 9.3), kernel="dirichlet", hp=True)
    for j in range(5):
        expected = ref_d.moments(M, vecs[j])
        assert rel_err(got[:, j], expected) < TOL[dtype]


def test_ldos_moments(setup):
    dtype, model, kpm, ref = setup
    fn = model.system.find_nearest
    idx = [fn([0, 0]), fn([1.0, 0.5], "B"), fn([-2, 1], "A"), fn([0.2, 0.1]), fn([3, 3])]
    M = 98
    for sel in (idx[:1], idx):
        got = kpm.impl.moments_ldos(M, sel)
        expected = ref.ldos_moments(M, sel)
        assert rel_err(got, expected) < TOL[dtype]


def test_greens_moments(setup):
    dtype, model, kpm, ref = setup
    n = model.system.num_sites
    i, j = n // 2, n // 4
    M = 98
    got = kpm.impl.moments_greens(M, i, [i])
    assert rel_err(got, ref.greens_moments(M, i, [i])) < TOL[dtype]
    cols = [j, j + 1, i, j + 7]
    got = kpm.impl.moments_greens(M, i, cols)
    expected = ref.greens_moments(M, i, cols)
    assert rel_err(got, expected) < TOL[dtype]


def test_kubo_moments(setup):
    dtype, model, kpm, ref = setup
    M = 34
    x, y = model.system.x, model.system.y
    for left, right in ((x, x), (x, y)):
        got = kpm.impl.moments_kubo(M, left, right, 2)
        expected = ref.kubo_moments(M, left, right, 2)
        assert rel_err(got, expected) < TOL[dtype] * 5


def test_curves(setup):
    dtype, model, kpm, ref = setup
    tol = TOL[dtype] * 10
    energy = np.linspace(-3, 3, 31)
    dos = kpm.calc_dos(energy, 0.3, num_random=4)
    assert rel_err(dos.data, ref.calc_dos(energy, 0.3, 4)) < tol
    i = model.system.find_nearest([0.5, 0.5])
    ldos = kpm.calc_ldos(energy, 0.3, [0.5, 0.5])
    assert rel_err(ldos.data, ref.calc_ldos(energy, 0.3, [i])[:, 0]) < tol
    n = model.system.num_sites
    g = kpm.calc_greens(n // 2, n // 3, energy, 0.3)
    assert rel_err(g, ref.calc_greens(n // 2, n // 3, energy, 0.3)) < tol
    gs = kpm.calc_greens(n // 2, [n // 3, n // 2], energy, 0.3)
    gr = ref.calc_greens(n // 2, [n // 3, n // 2], energy, 0.3)
    assert len(gs) == 2 and rel_err(gs[0], gr[0]) < tol and rel_err(gs[1], gr[1]) < tol


def test_conductivity_curve(setup):
    dtype, model, kpm, ref = setup
    mu = np.linspace(-2, 2, 11)
    sigma = kpm.calc_conductivity(mu, broadening=0.9, temperature=300, direction="xx", num_random=2, num_points=150)
    expected = ref.calc_conductivity(mu, 0.9, 300, model.system.x, model.system.x, num_random=2, num_points=150)
    assert rel_err(sigma.data, expected) < TOL[dtype] * 20
    sigma = kpm.calc_conductivity(mu, broadening=0.9, temperature=300, direction="xy", num_random=1, num_points=150)
    expected = ref.calc_conductivity(mu, 0.9, 300, model.system.x, model.system.y, num_random=1, num_points=150)
    assert np.abs(sigma.data - expected).max() < TOL[dtype] * 20 * max(np.abs(expected).max(), 1e-3)


def test_generic_moments(setup):
    """KPM.moments(alpha, beta, op): damped, truncated, mu_0 halved (Core.cpp:35-56)"""
    dtype, model, kpm, ref = setup
    rng = np.random.default_rng(1)
    n = model.hamiltonian.shape[0]
    cplx = dtype.kind == "c"
    alpha = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    beta = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    op = model.hamiltonian.copy()
    op.data = (op.data * rng.standard_normal(op.nnz)).astype(dtype)
    for args in ((alpha, None, None), (alpha, beta, None), (alpha, beta, op), (alpha, None, op)):
        got = kpm.moments(37, *args)
        expected = ref.moments(37, *args)
        assert got.shape == (37,)
        assert rel_err(got, expected) < TOL[dtype] * 5


def test_native_f32_reference_is_not_closer_than_gpu():
    """The reference's own f32 accumulation noise vs the double-accumulating answer, next to the GPU's distance"""
    model = pb.graphene_rectangle(30)  # 34 k sites
    kpm = pb.kpm(model, energy_range=(-8.5, 8.5), silent=True)
    M = 258
    hp = OracleKPM(model.hamiltonian, energy_range=(-8.5, 8.5), hp=True).dos_moments(M, 1)
    native = OracleKPM(model.hamiltonian, energy_range=(-8.5, 8.5), hp=False).dos_moments(M, 1)
    gpu = kpm.impl.moments_dos(M, 1)
    assert rel_err(gpu, hp) < 1e-5
    assert rel_err(gpu, hp) <= rel_err(native, hp) + 1e-7


def test_auto_bounds_lanczos():
    model = pb.graphene_rectangle(12)
    kpm = pb.kpm(model, silent=True)
    mn, mx, loops = kpm.impl.bounds
    b = OracleKPM(model.hamiltonian).bounds()
    assert mn == pytest.approx(b["min"], rel=1e-4) and mx == pytest.approx(b["max"], rel=1e-4)
    assert abs(loops - b["loops"]) <= 3
    a, _ = kpm.scaling_factors
    assert a == pytest.approx(b["a"], rel=1e-4)


@pytest.mark.parametrize("name,kw", [("graphene-pristine", dict(width=15)),
                                     ("graphene-const_potential", dict(width=15, onsite=0.5)),
                                     ("graphene-magnetic_field", dict(width=15, magnetic_field=1e3))])
def test_reference_golden_ldos(golden, name, kw):
    """The reference's own baseline curves (tests/test_kpm.py:25-47), auto bounds, both slicing modes"""
    model = pb.graphene_rectangle(**kw)
    energy = np.linspace(0, 2, 25)
    for optimal_size in (True, False):
        kpm = pb.kpm(model, kernel=pb.lorentz_kernel(), silent=True, optimal_size=optimal_size)
        ldos = kpm.calc_ldos(energy, broadening=0.15, position=[0, 0.07], reduce=False)
        assert np.allclose(ldos.data, golden["ldos[{}]".format(name)], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name,kw", [("graphene-const_potential", dict(width=25, onsite=0.5)),
                                     ("graphene-magnetic_field", dict(width=25, magnetic_field=1e3))])
def test_reference_golden_dos(golden, name, kw):
    model = pb.graphene_rectangle(**kw)
    kpm = pb.kpm(model, kernel=pb.lorentz_kernel(), silent=True)
    dos = kpm.calc_dos(np.linspace(0, 2, 25), broadening=0.15)
    assert np.allclose(dos.data, golden["dos[{}]".format(name)], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name,kw", [("graphene-const_potential", dict(width=20, onsite=0.5)),
                                     ("graphene-magnetic_field", dict(width=20, magnetic_field=1e3))])
def test_reference_golden_conductivity(golden, name, kw):
    model = pb.graphene_rectangle(**kw)
    kpm = pb.kpm(model, energy_range=[-9, 9], kernel=pb.lorentz_kernel(), silent=True)
    sigma = kpm.calc_conductivity(np.linspace(-2, 2, 25), broadening=0.5, temperature=0, num_points=200)
    assert np.allclose(sigma.data, golden["conductivity[{}]".format(name)], rtol=1e-2, atol=1e-5)


def test_api_contract():
    """Error strings / behaviours the reference's tests rely on (tests/test_kpm.py:50-114)"""
    model = pb.graphene_rectangle(6)
    kpm = pb.kpm(model, silent=True)
    with pytest.raises(RuntimeError) as excinfo:
        kpm.moments(10, [1, 2, 3])
    assert "Size mismatch" in str(excinfo.value)
    with pytest.raises(RuntimeError) as excinfo:
        kpm.moments(10, np.full(model.hamiltonian.shape[0], 1j))
    assert "Hamiltonian is real, but the given argument 'alpha' is complex" in str(excinfo.value)
    with pytest.raises(RuntimeError):
        kpm.calc_greens(-1, 0, [0.0], 0.1)
    with pytest.raises(RuntimeError):
        kpm.calc_conductivity([0.0], 0.5, 0, direction="xw")
    with pytest.raises(ValueError):
        pb.kpm(model, energy_range=(3, -3), silent=True)
    # reuse == fresh object (tests/test_kpm.py:104-114)
    energy = np.linspace(-5, 5, 50)
    for position in ([0, 0], [2, 0]):
        a = kpm.calc_ldos(energy, 0.1, position)
        b = pb.kpm(model, silent=True).calc_ldos(energy, 0.1, position)
        assert np.allclose(a.data, b.data, rtol=1e-3, atol=1e-6)
    # manual reconstruction from kpm.moments == calc_ldos (tests/test_kpm.py:50-77)
    idx = model.system.find_nearest([0, 0], "A")
    alpha = np.zeros(model.hamiltonian.shape[0])
    alpha[idx] = 1
    a, b = kpm.scaling_factors
    energy = np.linspace(0, 2, 25)
    num_moments = kpm.kernel.required_num_moments(0.15 / a)
    moments = kpm.moments(num_moments, alpha)
    ns = np.arange(num_moments)
    se = (energy - b) / a
    k = 2 / (a * np.pi * np.sqrt(1 - se**2))
    manual = k * np.sum(moments.real * np.cos(ns * np.arccos(se[:, np.newaxis])), axis=1)
    expected = kpm.calc_ldos(energy, 0.15, [0, 0], "A")
    assert np.allclose(manual, expected.data, rtol=1e-4, atol=1e-6)
    # deferred + progress + report + stats
    d = kpm.deferred_ldos(energy, 0.15, [0, 0])
    assert np.allclose(d.result.squeeze(), kpm.calc_ldos(energy, 0.15, [0, 0]).data)
    calls = []
    kp = pb.kpm(model, progress_callback=lambda delta, total: calls.append((delta, total)))
    kp.calc_dos(energy, 0.3, num_random=5)
    assert calls[0] == (-1, 5) and calls[-1] == (5, 5)
    s = kp.stats
    assert s.num_moments > 0 and s.eps > 0 and s.step_launches == s.num_moments // 2
    assert "moments" in kp.report() and "eps" in kp.report(True)
    sl = kpm.calc_spatial_ldos(energy, 0.3, pb.Rectangle(1.0))
    assert sl.data.shape == (25, len(sl.structure)) and len(sl.structure) > 10


def test_pybind11_binding_equals_ctypes_binding():
    """Both bindings drive the same C ABI: identical numbers, reference exception types, GIL released during the calls"""
    model = pb.graphene_rectangle(10, onsite=0.2, magnetic_field=300.0)
    energy = np.linspace(-2, 2, 41)
    a = pb.kpm(model, energy_range=(-9, 9), silent=True)
    calls = []
    b = pb.kpm(model, energy_range=[-9, 9], binding="pybind11", kernel=pb.jackson_kernel(),
               progress_callback=lambda delta, total: calls.append((delta, total)))
    assert a.scaling_factors == b.scaling_factors
    assert np.array_equal(a.calc_dos(energy, 0.2, num_random=3).data, b.calc_dos(energy, 0.2, num_random=3).data)
    assert calls[0] == (-1, 3) and calls[-1] == (3, 3)
    assert np.array_equal(a.calc_ldos(energy, 0.2, [0, 0]).data, b.calc_ldos(energy, 0.2, [0, 0]).data)
    assert np.array_equal(a.calc_ldos(energy, 0.2, [1, 1], "B", reduce=False).data, b.calc_ldos(energy, 0.2, [1, 1], "B", reduce=False).data)
    n = model.system.num_sites
    assert np.array_equal(a.calc_greens(n // 2, n // 3, energy, 0.2), b.calc_greens(n // 2, n // 3, energy, 0.2))
    ga, gb = a.calc_greens(5, [7, 9], energy, 0.2), b.calc_greens(5, [7, 9], energy, 0.2)
    assert len(gb) == 2 and all(np.array_equal(x, y) for x, y in zip(ga, gb))
    mu = np.linspace(-1, 1, 7)
    sa = a.calc_conductivity(mu, 0.8, 300, "xy", num_random=1, num_points=100)
    sb = b.calc_conductivity(mu, 0.8, 300, "xy", num_random=1, num_points=100)
    assert np.array_equal(sa.data, sb.data)
    alpha = np.zeros(model.hamiltonian.shape[0], np.complex128)
    alpha[3] = 1
    assert np.array_equal(a.moments(21, alpha), b.moments(21, alpha))
    assert np.array_equal(a.moments(21, alpha, alpha[::-1].copy(), model.hamiltonian),
                          b.moments(21, alpha, alpha[::-1].copy(), model.hamiltonian))
    sl_a = a.calc_spatial_ldos(energy, 0.3, pb.Rectangle(1.0))
    sl_b = b.calc_spatial_ldos(energy, 0.3, pb.Rectangle(1.0))
    assert np.array_equal(sl_a.data, sl_b.data)
    d = b.deferred_ldos(energy, 0.2, [0, 0])
    assert np.array_equal(np.asarray(d.result).squeeze(), a.calc_ldos(energy, 0.2, [0, 0]).data) and d.solver is b.impl
    assert "moments" in b.report() and b.stats.num_moments > 0 and b.stats.eps > 0
    assert b.kernel.required_num_moments(0.01) == a.kernel.required_num_moments(0.01)
    with pytest.raises(RuntimeError) as excinfo:
        b.moments(10, [1, 2, 3])
    assert "Size mismatch" in str(excinfo.value)
    with pytest.raises(RuntimeError):
        b.calc_greens(-1, 0, [0.0], 0.1)
    with pytest.raises(RuntimeError):
        b.calc_conductivity([0.0], 0.5, 0, direction="xw")
    with pytest.raises(ValueError):
        pb.kpm(model, energy_range=(3, -3), silent=True, binding="pybind11")
    b.model = pb.graphene_rectangle(8)     # cpb::KPM::set_model: new Hamiltonian, same object
    assert b.system.num_sites == pb.graphene_rectangle(8).system.num_sites
    assert np.array_equal(b.calc_dos(energy, 0.3, 2).data, pb.kpm(pb.graphene_rectangle(8), energy_range=(-9, 9), silent=True).calc_dos(energy, 0.3, 2).data)


# ---- models outside nearest-neighbour graphene: other ELL widths, several orbitals per site --------------------------
def test_reference_golden_ldos_mos2_three_orbitals(golden):
    """tests/test_kpm.py:23-47 for `group6_tmd.monolayer_3band("MoS2")`: LDOS per orbital (reduce=False) against the
    reference's baseline, and the orbital sum (reduce=True, KPM.cpp:64) -- ELL width 19, three Hamiltonian rows per site"""
    from pybinding_b200 import synthetic as syn
    model = syn.mos2_3band(pb.Rectangle(6))
    energy = np.linspace(0, 2, 25)
    for optimal_size in (True, False):
        for binding in ("ctypes", "pybind11"):
            kpm = pb.kpm(model, kernel=pb.lorentz_kernel(), silent=True, optimal_size=optimal_size, binding=binding)
            ldos = kpm.calc_ldos(energy, broadening=0.15, position=[0, 0.07], reduce=False)
            assert ldos.data.shape == (25, 3)
            assert np.allclose(ldos.data, golden["ldos[mos2]"], rtol=1e-3, atol=1e-6)
            total = kpm.calc_ldos(energy, broadening=0.15, position=[0, 0.07])
            assert total.data.shape == (25,) and np.allclose(total.data, ldos.data.sum(axis=1), rtol=1e-6)
    with pytest.raises(RuntimeError, match="multi-orbital"):
        kpm.calc_spatial_ldos(energy, 0.15, pb.Rectangle(1.0))
    # conductivity uses the positions expanded per orbital (KPM.cpp:140)
    sigma = pb.kpm(model, energy_range=(-4, 6), kernel=pb.lorentz_kernel(), silent=True) \
        .calc_conductivity(np.linspace(-1, 1, 5), broadening=0.5, temperature=300, num_random=1, num_points=60)
    ref = OracleKPM(model.hamiltonian, energy_range=(-4, 6), kernel="lorentz", hp=True)
    xs = model.system.expanded_positions.x
    expected = ref.calc_conductivity(np.linspace(-1, 1, 5), 0.5, 300, xs, xs, num_random=1, num_points=60)
    assert rel_err(sigma.data, expected) < 2e-4


@pytest.mark.parametrize("name", ["nnn_graphene_f32", "nnn_graphene_c128", "mos2_f32", "mos2_f64"])
def test_generic_ell_widths_match_the_oracle(name):
    """Lattices whose ELL width is not one of the nearest-neighbour cases (3, 4, 7): next-nearest-neighbour graphene
    (10 per row, 11 with the b offset) and the three-band TMD model (19 / 20); DOS, LDOS, Green's and generic moments"""
    from pybinding_b200 import synthetic as syn
    if name.startswith("nnn"):
        dtype = np.dtype(np.float32 if name.endswith("f32") else np.complex128)
        model = syn.graphene_monolayer(pb.Rectangle(9.0), nearest_neighbors=2, dtype=dtype,
                                       magnetic_field=300.0 if dtype.kind == "c" else 0.0)
        er = (-9.2, 9.6)
    else:
        dtype = np.dtype(np.float32 if name.endswith("f32") else np.float64)
        model = syn.mos2_3band(pb.Rectangle(14.0), dtype=dtype)
        er = (-3.5, 6.5)
    kpm = pb.kpm(model, energy_range=er, silent=True)
    ref = OracleKPM(model.hamiltonian, energy_range=er, hp=True)
    tol = TOL[dtype]
    for R in (1, 5, 16):
        assert rel_err(kpm.impl.moments_dos(130, R), ref.dos_moments(130, R)) < tol
    n = model.hamiltonian.shape[0]
    idx = [n // 2, n // 3, n // 2 + 1]
    assert rel_err(kpm.impl.moments_ldos(98, idx), ref.ldos_moments(98, idx)) < tol
    assert rel_err(kpm.impl.moments_ldos(98, idx[:1]), ref.ldos_moments(98, idx[:1])) < tol
    assert rel_err(kpm.impl.moments_greens(98, idx[0], idx), ref.greens_moments(98, idx[0], idx)) < tol
    x = model.system.expanded_positions.x
    assert rel_err(kpm.impl.moments_kubo(34, x, x, 1), ref.kubo_moments(34, x, x, 1)) < tol * 5
