"""GPU tests of the step-kernel variants and of size-independent properties of the moments.

* the bulk-copy (TMA) staged kernel `cheb_step_bulk` (kernels_bulk.cu) against the general kernel `cheb_step`
  on the same inputs, every dtype and ELL width, both x-staging modes and several pipeline depths, and against the
  CPU oracle (moment tolerance of north_star: 1e-5 for f32/c64, 1e-11 for f64/c128, relative to max |mu|);
* properties that hold at any size (used at benchmark-like sizes where the oracle would take too long):
  mu_0 = N / 2 exactly for the +-1 / unit-modulus stochastic starters, invariance under the batch size, the row
  order (locality tile) and the rank count, run-to-run bit reproducibility.

Engine tuning knobs are read from the environment when a context is created (engine.cu), so each variant is a
fresh `pb.kpm` object created under a patched environment.
"""
import os
from contextlib import contextmanager

import numpy as np
import pytest

import pybinding_b200 as pb
from oracle.oracle import OracleKPM

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5,
       np.dtype(np.float64): 1e-11, np.dtype(np.complex128): 1e-11}
KNOBS = ("PBK_PERSIST", "PBK_RES", "PBK_RES_TILE", "PBK_RES_ROW", "PBK_RES_CTAS", "PBK_RES_STAGES", "PBK_RES_BUFS", "PBK_RES_L2PF", "PBK_KUBO_WAVES", "PBK_KUBO_CHUNK", "PBK_RELEASE", "PBK_DEVBUILD", "PBK_BULK", "PBK_XS", "PBK_TILE", "PBK_TPB", "PBK_BPSM", "PBK_PF", "PBK_PFMASK", "PBK_MT_SEQUENTIAL",
         "PBK_CONE",
         "PBK_GRAPH", "PBK_GRAPH_MAX_MB")


@contextmanager
def knobs(**kw):
    saved = {k: os.environ.get(k) for k in KNOBS}
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def rel_err(actual, expected):
    actual, expected = np.asarray(actual), np.asarray(expected)
    return float(np.abs(actual - expected).max() / np.abs(expected).max())


def dos_moments(model, energy_range, M, R, max_batch=0, **kw):
    with knobs(**kw):
        kpm = pb.kpm(model, energy_range=energy_range, silent=True, max_batch=max_batch)
        mom = kpm.impl.moments_dos(M, R)
        return mom, kpm.stats


def model_for(dtype, k):
    dtype = np.dtype(dtype)
    if k == 7:
        if dtype.kind == "c":
            pytest.skip("the cubic generator is real")
        return pb.cubic_anderson(14, disorder=2.0, dtype=dtype), (-8.2, 8.2)   # 2744 sites, 6 neighbours + onsite
    field = 400.0 if dtype.kind == "c" else 0.0
    onsite = 0.3 if k == 4 else 0.0
    return pb.graphene_rectangle(9.0, onsite=onsite, dtype=dtype, magnetic_field=field), (-9, 9)


@pytest.mark.parametrize("k", [3, 4, 7])
@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_bulk_kernel_matches_general_kernel_and_oracle(dtype, k):
    dtype = np.dtype(dtype)
    model, er = model_for(dtype, k)
    M, R = 66, 12
    general, s0 = dos_moments(model, er, M, R, PBK_BULK=0)
    assert s0.bulk_launches == 0 and s0.step_launches == M // 2
    expected = OracleKPM(model.hamiltonian, energy_range=er, hp=True).dos_moments(M, R)
    assert rel_err(general, expected) < TOL[dtype]
    for kw in (dict(), dict(PBK_XS=0), dict(PBK_BULK=2), dict(PBK_BULK=7, PBK_XS=0), dict(PBK_BULK=3, PBK_TILE=64)):
        try:
            staged, s1 = dos_moments(model, er, M, R, **kw)
        except Exception as e:
            raise AssertionError("staged kernel failed with {}: {}".format(kw, e))
        assert s1.bulk_launches == M // 2 - 1, "the staged kernel did not run: {}".format(kw)  # all but the r1 = H r0 / 2 step
        assert rel_err(staged, expected) < TOL[dtype], kw
        # same arithmetic per row; only the per-thread partition of the f64 sums differs (f32 vectors: ~1e-8)
        assert rel_err(staged, general) < (1e-12 if dtype.itemsize >= 8 and dtype != np.complex64 else 1e-6), kw


@pytest.mark.parametrize("R", [1, 2, 5, 8, 33, 64])
def test_bulk_kernel_lane_counts(R):
    """Every chunks-per-row geometry: R = 1 falls back to the scalar general kernel, the others are staged"""
    model = pb.graphene_rectangle(12.0, dtype=np.complex64, magnetic_field=300.0)
    M = 34
    general, _ = dos_moments(model, (-9, 9), M, R, PBK_BULK=0)
    staged, s = dos_moments(model, (-9, 9), M, R)
    assert rel_err(staged, general) < 1e-6
    assert (s.bulk_launches > 0) == (R > 1)


def test_mu0_is_half_the_system_size():
    """mu_0 = <r|r> / 2 = N / 2 exactly for +-1 and for exp(i phi) starters, at any size"""
    for dtype, field in ((np.float32, 0.0), (np.complex64, 50.0)):
        model = pb.graphene_rectangle(120.0, dtype=dtype, magnetic_field=field)   # 0.55 M sites
        n = model.hamiltonian.shape[0]
        mom, s = dos_moments(model, (-8.5, 8.5), 18, 8)
        assert s.bulk_launches > 0
        assert abs(mom[0].real - n / 2) <= (0 if dtype == np.float32 else 2e-7 * n)
        assert abs(mom[0].imag) == 0
        # mu_1 = <r|H~|r>: purely real for a Hermitian H; |mu_n| <= N/2 * 2
        assert np.abs(mom.imag).max() <= 1e-6 * n
        assert np.abs(mom).max() <= n


def test_invariance_under_batching_order_and_reproducibility():
    model = pb.graphene_rectangle(60.0, dtype=np.complex64, magnetic_field=100.0)   # 137 k sites
    M, R = 130, 16
    base, s = dos_moments(model, (-8.5, 8.5), M, R)
    assert s.batch == 16 and s.bulk_launches > 0
    again, _ = dos_moments(model, (-8.5, 8.5), M, R)
    assert np.array_equal(base, again), "moments must be bit-reproducible run to run"
    scale = np.abs(base).max()
    for kw, mb in ((dict(), 4), (dict(), 6), (dict(PBK_TILE=-1), 0), (dict(PBK_TILE=1024), 0), (dict(PBK_MT_SEQUENTIAL=1), 0),
                   (dict(PBK_BULK=0, PBK_TILE=64), 0)):
        other, _ = dos_moments(model, (-8.5, 8.5), M, R, max_batch=mb, **kw)
        assert np.abs(other - base).max() / scale < 2e-6, (kw, mb)   # f32 vectors: only the summation order differs


def test_f64_invariance_is_tight():
    model = pb.graphene_rectangle(40.0, dtype=np.float64, onsite=0.2)
    M, R = 98, 6
    base, _ = dos_moments(model, (-9, 9), M, R)
    scale = np.abs(base).max()
    for kw, mb in ((dict(), 2), (dict(PBK_TILE=-1), 0), (dict(PBK_BULK=0), 0), (dict(PBK_XS=0, PBK_BULK=6), 3)):
        other, _ = dos_moments(model, (-9, 9), M, R, max_batch=mb, **kw)
        assert np.abs(other - base).max() / scale < 1e-12, (kw, mb)


def test_ldos_equals_dos_of_unit_vectors_and_sum_rule():
    """sum over all sites of the LDOS moments = trace moments: mu_n^{DOS-exact} = sum_i mu_n^{(i)} (small system)"""
    model = pb.graphene_rectangle(3.0, dtype=np.float64, onsite=0.1)
    n = model.hamiltonian.shape[0]
    kpm = pb.kpm(model, energy_range=(-9, 9), silent=True)
    M = 34
    ldos = kpm.impl.moments_ldos(M, list(range(n)))          # M x n
    h = model.hamiltonian.toarray().astype(np.float64)
    a, b = kpm.scaling_factors
    w = np.linalg.eigvalsh(h)
    t = np.cos(np.arange(M)[:, None] * np.arccos((w - b) / a)[None, :]).sum(axis=1)   # trace of T_n(H~)
    t[0] *= 0.5
    assert np.abs(ldos.sum(axis=1).real - t).max() / np.abs(t).max() < 1e-11


def test_spread_ldos_sites_use_full_system_layout_and_match_single_site_runs():
    """LDOS at sites spread over the sample (core.ldos(indices), cppcore/src/kpm/Core.cpp:58-72) without the light-cone
    sub-systems (PBK_CONE=0): the union of the light cones is the whole system, so the engine advances the unit vectors
    on the locality layout with the staged kernel; each column must equal the light-cone-sliced single-site run."""
    model = pb.graphene_rectangle(30.0, dtype=np.complex128, magnetic_field=200.0, disorder=0.3)
    fn = model.system.find_nearest
    sites = [fn([x, y]) for x in (-12, -4, 4, 12) for y in (-10, 0, 10)]
    M = 258
    with knobs(PBK_CONE=0):
        kpm = pb.kpm(model, energy_range=(-9.2, 9.2), silent=True)
        batch = kpm.impl.moments_ldos(M, sites)
        s = kpm.stats
        assert s.bulk_launches > 0 and s.opt_nnz == s.nnz          # full system, staged kernel
        single = pb.kpm(model, energy_range=(-9.2, 9.2), silent=True)
        for j, site in enumerate(sites):
            one = single.impl.moments_ldos(M, [site])[:, 0]
            assert np.abs(batch[:, j] - one).max() / np.abs(one).max() < 1e-11
        assert single.stats.bulk_launches == 0 and single.stats.opt_nnz < single.stats.nnz   # light-cone sliced
        # neighbouring sites (one cell) keep the sliced layout
        near = [fn([0, 0], "A"), fn([0, 0], "B")]
        both = single.impl.moments_ldos(M, near)
        assert single.stats.opt_nnz < single.stats.nnz
        for j, site in enumerate(near):
            one = kpm.impl.moments_ldos(M, [site])[:, 0]
            assert np.abs(both[:, j] - one).max() / np.abs(one).max() < 1e-11
    # default engine (light-cone sub-systems where they pay): same table
    default = pb.kpm(model, energy_range=(-9.2, 9.2), silent=True).impl.moments_ldos(M, sites)
    assert np.abs(default - batch).max() / np.abs(batch).max() < 1e-11


def ldos_moments(model, energy_range, M, sites, **kw):
    with knobs(**kw):
        kpm = pb.kpm(model, energy_range=energy_range, silent=True)
        return kpm.impl.moments_ldos(M, sites), kpm.stats


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_cone_ldos_equals_full_relabelling(dtype):
    """LDOS on the light-cone sub-system cut out of the resident Hamiltonian (moments_ldos_cones) against the
    reference's scheme (relabel the whole system from the source on the host, OptimizedHamiltonian.cpp:88-143):
    same slices, same rows, same moments."""
    dtype = np.dtype(dtype)
    model = pb.graphene_rectangle(25.0, dtype=dtype, onsite=0.2, magnetic_field=150.0 if dtype.kind == "c" else 0.0)
    site = model.system.find_nearest([1.0, -2.0])
    corner = model.system.find_nearest([-12.5, -12.5])
    for M in (34, 130, 514):        # light cone inside the sample, touching the edges, covering the whole sample
        for src in (site, corner):
            old, s0 = ldos_moments(model, (-9, 9), M, [src], PBK_CONE=0)
            new, s1 = ldos_moments(model, (-9, 9), M, [src])
            assert s1.opt_nnz == s0.opt_nnz and s1.nnz == s0.nnz, (M, src)
            assert s1.step_launches == s0.step_launches == M // 2
            assert rel_err(new, old) < (1e-12 if dtype in (np.float64, np.complex128) else 2e-6), (M, src)
    expected = OracleKPM(model.hamiltonian, energy_range=(-9, 9), hp=True).ldos_moments(130, [site])
    new, _ = ldos_moments(model, (-9, 9), 130, [site])
    assert rel_err(new, expected) < TOL[dtype]


def test_cone_ldos_many_sites_of_a_large_system():
    """Several sites of a large sample: each one runs on its own light-cone sub-system (no pass over the full system)"""
    model = pb.graphene_rectangle(200.0, dtype=np.float64, onsite=0.1)    # 1.5 M sites
    fn = model.system.find_nearest
    sites = [fn([-60, 10]), fn([0, 0]), fn([99.9, -99.9]), fn([35, 70])]
    M = 130
    batch, s = ldos_moments(model, (-9, 9), M, sites)
    assert s.num_batches == 1 and s.batch == len(sites) and s.bulk_launches == 0 and s.step_launches == M // 2   # one group
    assert s.opt_nnz < s.nnz / 50
    for j, site in enumerate(sites):
        one, s0 = ldos_moments(model, (-9, 9), M, [site], PBK_CONE=0)
        assert np.abs(batch[:, j] - one[:, 0]).max() / np.abs(one).max() < 1e-12
    assert np.all(batch[0].real == 0.5) and np.all(batch.imag == 0)


def test_graph_replay_of_small_recursions_is_bit_identical():
    """Launch-bound recursions (small systems) are captured once as a CUDA graph and replayed: same kernels, same
    parameters, so the moments are bit-identical to plain launches, call after call, for DOS and for the light-cone
    sliced recursion on the relabelled Hamiltonian (PBK_CONE=0).  (A single vector normally runs in the persistent
    kernel instead: PBK_PERSIST=0 selects the launch-per-step path here.)"""
    model = pb.graphene_rectangle(40.0, dtype=np.float32)     # configs[0]: 61 k sites
    M = 1026
    site = model.system.find_nearest([3, 4])
    with knobs(PBK_GRAPH=0, PBK_PERSIST=0, PBK_CONE=0):
        plain_kpm = pb.kpm(model, energy_range=(-8.5, 8.5), silent=True)
        plain = plain_kpm.impl.moments_dos(M, 1)
        assert plain_kpm.stats.graph_launches == 0 and plain_kpm.stats.persist_launches == 0
        plain_ldos = pb.kpm(model, energy_range=(-8.5, 8.5), silent=True).impl.moments_greens(M, site, [site])
    with knobs(PBK_PERSIST=0, PBK_CONE=0):
        kpm = pb.kpm(model, energy_range=(-8.5, 8.5), silent=True)
        for call in range(3):
            mom = kpm.impl.moments_dos(M, 1)
            s = kpm.stats
            assert s.graph_launches == 1 and s.step_launches == M // 2 and s.kernel_launches > M // 2
            assert np.array_equal(mom, plain), call
        for call in range(2):   # light-cone sliced diagonal Green's function through the relabelled Hamiltonian
            g = kpm.impl.moments_greens(M, site, [site])
            assert kpm.stats.graph_launches == 1
            assert np.array_equal(g, plain_ldos), call
    # the same diagonal element on the light-cone sub-system (default path) carries the same numbers
    # (float32: the rows keep the slot order of the resident layout there, so the FMAs round differently)
    cone = pb.kpm(model, energy_range=(-8.5, 8.5), silent=True).impl.moments_greens(M, site, [site])
    assert rel_err(cone, plain_ldos) < 2e-6
    big = pb.graphene_rectangle(60.0, dtype=np.complex64, magnetic_field=100.0)
    with knobs(PBK_GRAPH_MAX_MB=1):
        k2 = pb.kpm(big, energy_range=(-8.5, 8.5), silent=True)
        k2.impl.moments_dos(66, 16)
        assert k2.stats.graph_launches == 0      # 137 k sites x 16 vectors x 8 bytes = 17.6 MB > 1 MB: plain launches


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_layouts_built_on_the_device_equal_the_host_build(dtype):
    """build.cu (CSR -> scaled, relabelled ELL by one kernel) against the host restatement of create_scaled /
    create_reordered (bit-identical to the oracle: tests/test_host_ell.py): the same matrix bits give the same moments
    bit for bit -- full-system locality layout with b != 0 (inserted diagonal), breadth-first layout (Green's), unscaled
    layout (Lanczos bounds) and the velocity operators (Kubo-Bastin)"""
    dtype = np.dtype(dtype)
    model = pb.graphene_rectangle(9.0, onsite=0.3 if dtype != np.float32 else 0.0, dtype=dtype,
                                  magnetic_field=400.0 if dtype.kind == "c" else 0.0)
    n = model.system.num_sites
    results = []
    for dev in (1, 0):
        with knobs(PBK_DEVBUILD=dev, PBK_CONE=0):
            kpm = pb.kpm(model, energy_range=(-9.1, 9.3), silent=True)
            auto = pb.kpm(model, silent=True)
            results.append(dict(dos=kpm.impl.moments_dos(66, 5), greens=kpm.impl.moments_greens(50, n // 2, [n // 3, n // 2 + 3]),
                                ldos=kpm.impl.moments_ldos(50, [n // 4]), bounds=auto.impl.bounds,
                                kubo=kpm.impl.moments_kubo(18, model.system.x, model.system.y, 1)))
    for key in results[0]:
        assert np.array_equal(np.asarray(results[0][key]), np.asarray(results[1][key])), key


@pytest.mark.parametrize("k", [3, 4, 7, 10])
@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_resident_tile_kernel_matches_general_kernel_and_oracle(dtype, k):
    """`cheb_step_res` (kernels_res.cu: x rows of a tile and its halo resident in shared memory, 16-bit local codes)
    against the general kernel and the oracle: every dtype, the specialised widths and a generic one, tiles that have
    to be split because their halo does not fit, several row widths / pipeline depths / CTAs per SM"""
    dtype = np.dtype(dtype)
    if k == 10:
        from pybinding_b200 import synthetic as syn
        model = syn.graphene_monolayer(pb.Rectangle(9.0), nearest_neighbors=2, dtype=dtype, magnetic_field=300.0 if dtype.kind == "c" else 0.0)
        er = (-9.6, 9.6)
    else:
        model, er = model_for(dtype, k)
    M = 66
    lanes = 64 // dtype.itemsize                     # one pass of the default 64-byte rows
    R = 2 * lanes + 1                                # two full passes and a ragged one
    general, s0 = dos_moments(model, er, M, R, PBK_BULK=0, PBK_RES=0)
    assert s0.res_launches == 0
    expected = OracleKPM(model.hamiltonian, energy_range=er, hp=True).dos_moments(M, R)
    assert rel_err(general, expected) < TOL[dtype]
    for kw in (dict(PBK_RES_TILE=128), dict(PBK_RES_TILE=64, PBK_RES_STAGES=4, PBK_RES_CTAS=1),
               dict(PBK_RES_TILE=256, PBK_RES_ROW=128, PBK_RES_CTAS=2), dict(PBK_RES_TILE=1024, PBK_RES_ROW=32, PBK_RES_CTAS=3)):
        res, s1 = dos_moments(model, er, M, R, PBK_RES=2, **kw)
        width = kw.get("PBK_RES_ROW", 64) // dtype.itemsize
        assert s1.batch == width and s1.num_batches == -(-R // width)
        # every pass (also the ragged one, padded with zero lanes to the full row width) runs the resident-tile kernel
        assert s1.res_launches == (M // 2 - 1) * s1.num_batches, "the resident-tile kernel did not run: {} {}".format(kw, s1.res_launches)
        assert rel_err(res, expected) < TOL[dtype], kw
        assert rel_err(res, general) < (1e-12 if dtype.itemsize >= 8 and dtype != np.complex64 else 1e-6), kw
    again, _ = dos_moments(model, er, M, R, PBK_RES=2, PBK_RES_TILE=128)
    first, _ = dos_moments(model, er, M, R, PBK_RES=2, PBK_RES_TILE=128)
    assert np.array_equal(again, first), "moments must be bit-reproducible run to run"


@pytest.mark.parametrize("k", [3, 4, 7])
@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_persistent_kernel_matches_launch_per_step_and_oracle(dtype, k):
    """`cheb_persistent` (kernels_persist.cu): the whole diagonal recursion of one vector in one cooperative launch with a
    grid barrier per step, against one launch per step and the oracle; one, two and four rows per thread"""
    dtype = np.dtype(dtype)
    model, er = model_for(dtype, k)
    M = 130
    stepwise, s0 = dos_moments(model, er, M, 1, PBK_PERSIST=0)
    assert s0.persist_launches == 0
    single, s1 = dos_moments(model, er, M, 1)
    assert s1.persist_launches == 1 and s1.step_launches == M // 2 and s1.graph_launches == 0
    expected = OracleKPM(model.hamiltonian, energy_range=er, hp=True).dos_moments(M, 1)
    assert rel_err(single, expected) < TOL[dtype]
    assert rel_err(single, stepwise) < (1e-12 if dtype.itemsize >= 8 and dtype != np.complex64 else 1e-6)
    again, _ = dos_moments(model, er, M, 1)
    assert np.array_equal(single, again), "moments must be bit-reproducible run to run"
    if k == 3 and dtype == np.float32:
        for width, rows_per_thread in ((55.0, 2), (85.0, 4)):     # 115 k / 276 k sites: more than one row per thread
            big = pb.graphene_rectangle(width, dtype=dtype)
            a, sa = dos_moments(big, er, 66, 1)
            b, sb = dos_moments(big, er, 66, 1, PBK_PERSIST=0)
            assert sa.persist_launches == 1 and sb.persist_launches == 0, rows_per_thread
            assert rel_err(a, b) < 1e-6
        huge = pb.graphene_rectangle(120.0, dtype=dtype)          # 550 k sites: too many rows for one resident grid
        _, sh = dos_moments(huge, er, 34, 1)
        assert sh.persist_launches == 0


def test_resident_tile_kernel_with_more_tiles_than_the_descriptor_window():
    """A CTA of `cheb_step_res` keeps a window of 128 tile descriptors in shared memory and reloads it as it goes; small
    tiles on a system of 1.4 M sites give every CTA ~150 tiles, so the reload and the look-ups past the window (producer,
    tile prefetch) are exercised.  Checked against the staged kernel on the same starters."""
    model = pb.graphene_rectangle(190.0, dtype=np.float32, onsite=0.1)
    er = (-9.0, 9.2)
    M, R = 18, 16
    staged, s0 = dos_moments(model, er, M, R, PBK_RES=0)
    assert s0.res_launches == 0
    res, s1 = dos_moments(model, er, M, R, PBK_RES=2, PBK_RES_TILE=64, PBK_RES_CTAS=1)
    n = model.hamiltonian.shape[0]
    assert n / 64 > 148 * 128, "the system is too small to overflow the descriptor window"
    assert s1.res_launches == M // 2 - 1
    assert rel_err(res, staged) < 1e-6
    assert abs(res[0].real - n / 2) < 1e-6 * n


@pytest.mark.parametrize("dtype", [np.float64, np.complex64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("M", [18, 128, 130, 134, 136, 200, 262])
def test_kubo_gemm_tile_shapes(dtype, M):
    """The Kubo-Bastin contraction (kubo.cu) for every tile situation of the 128 x 128 tiling: a single edge tile (18), one
    full tile (128), remainders folded into the last tiles (130, 134, 262 = 2 x 128 + 6), a remainder that gets its own tile
    row (136, 200) -- real and complex stacks, against the hp oracle."""
    dtype = np.dtype(dtype)
    model = pb.graphene_rectangle(6.0, dtype=dtype, onsite=0.2, magnetic_field=300.0 if dtype.kind == "c" else 0.0)
    er = (-9.0, 9.4)
    kpm = pb.kpm(model, energy_range=er, silent=True)
    ref = OracleKPM(model.hamiltonian, energy_range=er, hp=True)
    x, y = model.system.x, model.system.y
    got = kpm.impl.moments_kubo(M, x, y, 3)          # three vectors: two lanes + a ragged one for f64, padded lanes for c64
    expected = ref.kubo_moments(M, x, y, 3)
    assert got.shape == (M, M)
    assert rel_err(got, expected) < TOL[dtype] * 5
