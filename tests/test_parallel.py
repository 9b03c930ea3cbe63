"""Sweep layer (pybinding_b200/parallel.py): the contract of `_pybinding.parallel_for` (cppmodule/src/parallel.cpp:15-45)
and of `pybinding.parallel.{parallelize, parallel_for, sweep, ndsweep}` -- modelled on the reference's
tests/test_parallel.py:16-52 (sequential == parallel, results in sequence order)."""
import threading
import time

import numpy as np
import pytest

from pybinding_b200 import parallel
from pybinding_b200.chebyshev import Deferred
from pybinding_b200.results import Series


class Probe:
    def __init__(self):
        self.lock = threading.Lock()
        self.running = 0
        self.peak = 0

    def job(self, value, seconds=0.02):
        def compute():
            with self.lock:
                self.running += 1
                self.peak = max(self.peak, self.running)
            time.sleep(seconds)
            with self.lock:
                self.running -= 1
            return value * value
        return Deferred(None, compute)


def test_parallel_for_contract():
    probe = Probe()
    produced, retired = [], {}
    main = threading.get_ident()

    def produce(var):
        assert threading.get_ident() == main     # the producer runs on the calling thread, in sequence order
        produced.append(var)
        return probe.job(var)

    def retire(deferred, idx):
        assert threading.get_ident() == main
        retired[idx] = deferred.result

    seq = list(range(23))
    parallel._parallel_for(seq, produce, retire, num_threads=4, queue_size=2)
    assert produced == seq
    assert retired == {i: i * i for i in seq}
    assert 2 <= probe.peak <= 4                  # jobs overlap, never more than num_threads at once
    # single-threaded path == sequential loop
    retired.clear()
    parallel._parallel_for(seq, produce, retire, num_threads=1)
    assert retired == {i: i * i for i in seq}


def test_parallel_for_propagates_errors():
    def produce(var):
        def compute():
            if var == 5:
                raise RuntimeError("job 5 failed")
            return var
        return Deferred(None, compute)

    with pytest.raises(RuntimeError, match="job 5 failed"):
        parallel._parallel_for(range(12), produce, lambda d, i: None, num_threads=3)
    with pytest.raises(ValueError, match="bad produce"):
        parallel._parallel_for(range(4), lambda v: (_ for _ in ()).throw(ValueError("bad produce")), lambda d, i: None,
                               num_threads=2)


def test_parallelize_sweep_and_ndsweep():
    x = np.linspace(0, 1, 7)

    @parallel.parallelize(num_threads=3, v=x)
    def factory(v):
        energy = np.linspace(-1, 1, 5)
        return Deferred(None, lambda: Series(energy, v * energy))

    result = parallel.sweep(factory, labels=dict(x="v"))
    assert result.data.shape == (7, 5) and np.allclose(result.x, x) and np.allclose(result.y, np.linspace(-1, 1, 5))
    assert np.allclose(result.data, np.outer(x, np.linspace(-1, 1, 5)))

    @parallel.parallelize(num_threads=2, a=[1, 2, 3], b=[10, 20])
    def grid(a, b):
        return Deferred(None, lambda: np.array([a + b, a * b]))

    nd = parallel.ndsweep(grid)
    assert nd.data.shape == (3, 2, 2)
    assert nd.data[2, 1].tolist() == [23, 60]
    assert parallel.parallel_for(grid)[3].tolist() == [22, 40]    # (a, b) = (2, 20): product order


@pytest.mark.gpu
def test_deferred_ldos_sweep_on_the_gpu_equals_sequential():
    """Jobs on distinct KPM objects (own context, own stream) run concurrently and give the sequential results"""
    import pybinding_b200 as pb
    model = pb.graphene_rectangle(20.0, dtype=np.float64, onsite=0.1)
    energy = np.linspace(-2, 2, 41)
    xs = np.linspace(-8, 8, 9)
    ndev = max(1, min(2, parallel.num_devices()))      # two devices are enough to exercise the round-robin

    @parallel.parallelize(num_threads=4, queue_size=4, devices=list(range(ndev)), x=xs)
    def factory(x):
        kpm = pb.kpm(model, energy_range=(-9, 9), silent=True, device=parallel.device_for())
        return kpm.deferred_ldos(energy, broadening=0.1, position=[x, 0.5])

    result = parallel.sweep(factory)
    assert result.data.shape == (len(xs), len(energy))
    one = pb.kpm(model, energy_range=(-9, 9), silent=True)
    for i, x in enumerate(xs):
        expected = one.calc_ldos(energy, broadening=0.1, position=[x, 0.5]).data
        assert np.array_equal(result.data[i], expected)


@pytest.mark.gpu
def test_reference_sweep_goldens_on_the_gpu():
    """The reference's own sweep tests (tests/test_parallel.py:16-52) against its own baselines
    (tests/baseline_data/parallel/{sweep,ndsweep}.pbz -> tests/golden/reference_parallel_baselines.npz): deferred LDOS of
    a graphene armchair hexagon in a constant potential, through parallelize / sweep / ndsweep, reference tolerances."""
    import os
    import pybinding_b200 as pb
    from pybinding_b200 import synthetic as syn
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_parallel_baselines.npz"))
    shape = syn.graphene_hexagon_ac(side_width=15)

    @parallel.parallelize(v=np.linspace(0, 0.1, 10))
    def factory(v, energy=np.linspace(0, 0.1, 10)):
        model = syn.graphene_monolayer(shape, onsite=v)
        kpm = pb.kpm(model, kernel=pb.lorentz_kernel(), silent=True, device=parallel.device_for())
        return kpm.deferred_ldos(energy, broadening=0.15, position=[0, 0], sublattice="B")

    labels = dict(title="test sweep", x="V (eV)", y="E (eV)", data="LDOS")
    result = parallel.sweep(factory, labels=labels)
    assert np.allclose(result.x, golden["sweep.x"]) and np.allclose(result.y, golden["sweep.y"])
    assert np.allclose(result.data, golden["sweep.data"], rtol=1e-3, atol=1e-6)

    @parallel.parallelize(v1=np.linspace(0, 0.1, 5), v2=np.linspace(-0.2, 0.2, 4))
    def factory2(v1, v2, energy=np.linspace(0, 0.1, 10)):
        # pb.constant_potential(v1) then pb.constant_potential(v2): two float32 additions to the onsite energy
        model = syn.graphene_monolayer(shape, onsite=float(np.float32(v1) + np.float32(v2)))
        kpm = pb.kpm(model, kernel=pb.lorentz_kernel(), silent=True, device=parallel.device_for())
        return kpm.deferred_ldos(energy, broadening=0.15, position=[0, 0])

    nd = parallel.ndsweep(factory2)
    assert nd.data.shape == golden["ndsweep.data"].shape
    assert np.allclose(nd.data, golden["ndsweep.data"], rtol=1e-3, atol=1e-6)
