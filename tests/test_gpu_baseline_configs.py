"""GPU parity at the sizes BASELINE.json names: every configuration is compared with the oracle at its OWN size.

configs[0] is compared in full (61 k sites, float32, 1026 moments, 1 vector).  For configs[1] and configs[4] the full
Hamiltonian and the real MT19937 starters are used and the first 10 moments of the benched instantiation (4 and 64
lanes per pass) are compared -- the recursion is the same kernel launch sequence for every later moment.  configs[2]
is one LDOS site and one Green's pair at the configuration's own number of moments on the 9.5 M-site system,
configs[3] the Kubo-Bastin moment matrix at 200 nm with 66 moments.  Tolerances (north_star): 1e-5 (f32 / c64) and
1e-11 (f64 / c128) of max |mu|, against the `hp` oracle (f64 accumulation); the distance of the reference-faithful
`native` oracle to the same `hp` result is reported next to it (gpurun_out/parity_at_size.jsonl when writable).
"""
import json
import os
import time

import numpy as np
import pytest

import pybinding_b200 as pb
from oracle.oracle import OracleKPM, hardware_threads

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5,
       np.dtype(np.float64): 1e-11, np.dtype(np.complex128): 1e-11}


def rel_err(actual, expected):
    actual, expected = np.asarray(actual), np.asarray(expected)
    return float(np.abs(actual - expected).max() / np.abs(expected).max())


def record(**kw):
    """One line per comparison, for profiles/ (best effort: the driver's box may be read-only)"""
    print("PARITY", json.dumps(kw))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_at_size.jsonl"), "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def test_config0_graphene_40nm_f32_full_length():
    """configs[0] exactly: 40 x 40 nm, float32, 1026 moments, one random vector"""
    model = pb.graphene_rectangle(40.0)
    er = (-8.5, 8.5)
    M = 1026
    kpm = pb.kpm(model, energy_range=er, silent=True)
    gpu = kpm.impl.moments_dos(M, 1)
    assert kpm.stats.num_moments == M
    hp = OracleKPM(model.hamiltonian, energy_range=er, hp=True).dos_moments(M, 1)
    native = OracleKPM(model.hamiltonian, energy_range=er, hp=False).dos_moments(M, 1)
    e_hp, e_native = rel_err(gpu, hp), rel_err(native, hp)
    record(config=0, sites=int(model.hamiltonian.shape[0]), dtype="float32", moments=M, vectors=1,
           gpu_vs_hp=e_hp, native_vs_hp=e_native, gpu_vs_native=rel_err(gpu, native))
    assert e_hp < TOL[np.dtype(np.float32)]
    # the DOS curve of the same run, through the public API
    energy = np.linspace(-3, 3, 200)
    a = kpm.scaling_factors[0]
    broadening = float(np.float32(np.pi)) * a / (M - 2)
    dos = kpm.calc_dos(energy, broadening, num_random=1)
    assert kpm.stats.num_moments == M
    expected = OracleKPM(model.hamiltonian, energy_range=er, hp=True).calc_dos(energy, broadening, 1)
    assert rel_err(dos.data, expected) < 1e-4


@pytest.mark.parametrize("name", ["config1_graphene_1000nm_c64", "config4_cubic_256_f32"])
def test_full_size_dos_first_moments(name):
    """Full Hamiltonian, real MT19937 starters, first 10 moments with 4 and with 64 lanes per pass"""
    if name.startswith("config1"):
        model = pb.graphene_rectangle(1000.0, magnetic_field=10.0, dtype=np.complex64)
        er, cfg = (-8.5, 8.5), 1
    else:
        model = pb.cubic_anderson(256, disorder=4.0, seed=0, dtype=np.float32)
        er, cfg = (-8.2, 8.2), 4
    dtype = np.dtype(model.hamiltonian.dtype)
    n = model.hamiltonian.shape[0]
    M = 10
    threads = hardware_threads()
    kpm = pb.kpm(model, energy_range=er, silent=True)
    ref = OracleKPM(model.hamiltonian, energy_range=er, hp=True, num_threads=threads)
    for R in (4, 64):
        t0 = time.time()
        gpu = kpm.impl.moments_dos(M, R)
        s = kpm.stats
        # the benched kernels must be the ones that ran: the staged kernel (graphene) or the resident-tile kernel (cubic:
        # 16 float lanes per pass whatever the total), for every step after r1 = H r0 / 2
        assert s.bulk_launches + s.res_launches == (M // 2 - 1) * s.num_batches, (s.bulk_launches, s.res_launches, s.num_batches)
        assert s.batch == (R if s.res_launches == 0 else min(R, 16))
        t1 = time.time()
        hp = ref.dos_moments(M, R)
        e_hp = rel_err(gpu, hp)
        row = dict(config=cfg, sites=int(n), dtype=dtype.name, moments=M, vectors=R, gpu_vs_hp=e_hp,
                   gpu_seconds=round(t1 - t0, 2), oracle_seconds=round(time.time() - t1, 2))
        if R == 4:
            native = OracleKPM(model.hamiltonian, energy_range=er, hp=False, num_threads=threads).dos_moments(M, R)
            row.update(native_vs_hp=rel_err(native, hp), gpu_vs_native=rel_err(gpu, native))
        record(**row)
        assert e_hp < TOL[dtype], row
        # mu_0 = N / 2: exactly for the +-1 starters, to float32 rounding of |exp(i phi)|^2 for the complex ones
        assert gpu[0].real == pytest.approx(n / 2, rel=1e-12 if dtype.kind == "f" else 1e-7)


def test_config2_ldos_and_greens_at_size():
    """configs[2]: 500 x 500 nm, onsite disorder + Peierls field, complex128: one LDOS site and one Green's pair at the
    configuration's own number of moments (broadening 0.02 eV)"""
    model = pb.graphene_rectangle(500.0, disorder=0.5, disorder_seed=0, magnetic_field=10.0, dtype=np.complex128)
    er = (-8.6, 8.6)
    kpm = pb.kpm(model, energy_range=er, silent=True)
    ref = OracleKPM(model.hamiltonian, energy_range=er, hp=True)
    a = kpm.scaling_factors[0]
    M = kpm.kernel.required_num_moments(0.02 / a)
    assert M > 1300
    i = model.system.find_nearest([3.0, -2.0], "A")
    j = model.system.find_nearest([4.1, -1.3], "B")
    gpu = kpm.impl.moments_ldos(M, [i])
    s = kpm.stats
    assert not s.uses_full_system and s.opt_nnz < 0.1 * s.nnz     # the light cone never reaches the edge
    e_ldos = rel_err(gpu, ref.ldos_moments(M, [i]))
    gpu_g = kpm.impl.moments_greens(M, i, [j])
    e_greens = rel_err(gpu_g, ref.greens_moments(M, i, [j]))
    record(config=2, sites=int(model.hamiltonian.shape[0]), dtype="complex128", moments=int(M), ldos_vs_hp=e_ldos,
           greens_vs_hp=e_greens)
    assert e_ldos < 1e-11 and e_greens < 1e-11


def test_config3_kubo_moments_200nm():
    """configs[3]: 200 x 200 nm, float64, Kubo-Bastin moment matrix (xx and xy) with 66 moments"""
    model = pb.graphene_rectangle(200.0, dtype=np.float64)
    er = (-9.0, 9.0)
    kpm = pb.kpm(model, energy_range=er, kernel=pb.lorentz_kernel(), silent=True)
    ref = OracleKPM(model.hamiltonian, energy_range=er, kernel="lorentz", hp=True, num_threads=hardware_threads())
    x, y = model.system.x, model.system.y
    M = 66
    for name, (left, right) in (("xx", (x, x)), ("xy", (x, y))):
        gpu = kpm.impl.moments_kubo(M, left, right, 2)
        expected = ref.kubo_moments(M, left, right, 2)
        err = rel_err(gpu, expected)
        record(config=3, sites=int(model.hamiltonian.shape[0]), dtype="float64", moments=M, direction=name, vectors=2,
               gpu_vs_hp=err)
        assert err < 5e-11
