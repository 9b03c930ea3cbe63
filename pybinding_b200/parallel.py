"""Multi-stream / multi-GPU parameter sweeps over `Deferred` KPM jobs.

Counterpart of the reference's sweep layer for the KPM path:
  * `_pybinding.parallel_for(sequence, produce, retire, num_threads, queue_size)` (cppmodule/src/parallel.cpp:15-45,
    cppcore/include/detail/thread.hpp:167-201): a producer makes `Deferred` jobs in sequence order, `num_threads`
    workers call `compute()` concurrently, `retire(job, idx)` is called for every finished job;
  * `pybinding.parallel.{parallel_for, parallelize, sweep, ndsweep}` (pybinding/parallel.py:22-70,143-430), without the
    plotting / progress-bar / file-saving hooks, which are not part of the KPM path.

B200 mapping: a worker is a host thread; every job owns its own `pbk_ctx`, i.e. its own CUDA stream and buffers
(include/pbkpm.h: distinct contexts are fully concurrent), so jobs overlap on one GPU where they are launch- or
host-bound and spread over several GPUs when `devices` names more than one (`device_for(idx)` tells `produce` which
device job `idx` should build its solver on).  ctypes / pybind11 release the GIL during every engine call, like
`wrappers.hpp:15` does for the reference.
"""
import inspect
import itertools
import queue
import threading

import numpy as np

__all__ = ["num_devices", "device_for", "parallel_for", "parallelize", "sweep", "ndsweep", "Sweep", "NDSweep"]

_context = threading.local()


def num_devices():
    """CUDA devices visible to the engine (0 without a usable device)"""
    import ctypes
    from . import _lib
    count = ctypes.c_int(0)
    status = _lib.load().pbk_device_count(ctypes.byref(count))
    return count.value if status == 0 else 0


def device_for(idx=None):
    """Device the job being produced should run on (round-robin over `devices` of the running `parallel_for`)"""
    devices = getattr(_context, "devices", None) or [0]
    if idx is None:
        idx = getattr(_context, "index", 0)
    return devices[idx % len(devices)]


def _sequential_for(sequence, produce, retire):
    for idx, var in enumerate(sequence):
        _context.index = idx
        deferred = produce(var)
        deferred.compute()
        retire(deferred, idx)


def _parallel_for(sequence, produce, retire, num_threads=4, queue_size=None, devices=None):
    """Same contract as `_pybinding.parallel_for`: `produce(var) -> Deferred` is called in sequence order (on the
    calling thread, like the reference's GIL-holding producer), at most `queue_size` produced jobs wait for a worker,
    `num_threads` workers run `compute()`, `retire(deferred, idx)` is called on the calling thread for every finished
    job.  The first exception raised by any stage is re-raised after the workers have stopped."""
    sequence = list(sequence)
    num_threads = max(1, int(num_threads))
    queue_size = max(1, int(queue_size if queue_size is not None else num_threads))
    _context.devices = list(devices) if devices else [0]
    if num_threads == 1:
        return _sequential_for(sequence, produce, retire)

    jobs = queue.Queue(maxsize=queue_size)
    done = queue.Queue()
    errors = []
    stop = threading.Event()

    def worker():
        while True:
            item = jobs.get()
            if item is None:
                return
            idx, deferred = item
            try:
                if not stop.is_set():
                    deferred.compute()
            except BaseException as exc:  # noqa: B902 -- reported to the caller below
                errors.append(exc)
                stop.set()
            done.put((idx, deferred))

    threads = [threading.Thread(target=worker, daemon=True) for _ in range(num_threads)]
    for t in threads:
        t.start()
    retired = 0

    def drain(block):
        nonlocal retired
        while True:
            try:
                idx, deferred = done.get(block=block and retired < produced, timeout=None if block else 0)
            except queue.Empty:
                return
            retired += 1
            if not stop.is_set():
                retire(deferred, idx)
            if block and retired >= produced:
                return

    produced = 0
    try:
        for idx, var in enumerate(sequence):
            if stop.is_set():
                break
            _context.index = idx
            deferred = produce(var)
            jobs.put((idx, deferred))
            produced += 1
            drain(block=False)
    except BaseException as exc:  # noqa: B902
        errors.append(exc)
        stop.set()
    finally:
        for _ in threads:
            jobs.put(None)
        try:
            if retired < produced:
                drain(block=True)
        except BaseException as exc:  # noqa: B902
            errors.append(exc)
        for t in threads:
            t.join()
    if errors:
        raise errors[0]


class Sweep:
    """x, y, data container of `sweep` (reference: pybinding/results.py `Sweep`)"""

    def __init__(self, x, y, data, labels=None, tags=None):
        self.x, self.y, self.data = np.asarray(x), np.asarray(y), np.asarray(data)
        self.labels = dict(labels or {})
        self.tags = dict(tags or {})


class NDSweep:
    """variables, data container of `ndsweep` (reference: pybinding/results.py `NDSweep`)"""

    def __init__(self, variables, data, labels=None, tags=None):
        self.variables = [np.asarray(v) for v in variables]
        self.data = np.asarray(data)
        self.labels = dict(labels or {})
        self.tags = dict(tags or {})


class _Factory:
    def __init__(self, variables, produce, num_threads, queue_size, devices, fixtures=None):
        self.variables = [np.atleast_1d(v) for v in variables]
        self.sequence = list(itertools.product(*self.variables))
        self.produce = produce
        self.fixtures = dict(fixtures or {})   # keyword defaults of the factory function, e.g. `energy` (parallel.py:352-353)
        self.num_threads, self.queue_size, self.devices = num_threads, queue_size, devices


def parallelize(num_threads=4, queue_size=None, devices=None, **variables):
    """Decorator: `@parallelize(a=values_a, b=values_b) def factory(a, b): return kpm.deferred_ldos(...)`
    (reference: pybinding/parallel.py:314-360); the product of the keyword sequences is the sweep."""
    def decorator(produce_func):
        params = inspect.signature(produce_func).parameters
        names = [k for k in params if k in variables]          # the order of the function's parameters, like the reference
        fixtures = {k: v.default for k, v in params.items() if k not in variables and v.default is not inspect.Parameter.empty}

        def produce(values):
            return produce_func(**dict(zip(names, values)))
        return _Factory([variables[n] for n in names], produce, num_threads, queue_size, devices, fixtures)
    return decorator


def parallel_for(factory, make_result=None):
    """Run a `parallelize`d factory; returns the list of results in sequence order, or `make_result(list)`
    (reference: pybinding/parallel.py:282-311)"""
    results = [None] * len(factory.sequence)

    def retire(deferred, idx):
        results[idx] = deferred.result

    _parallel_for(factory.sequence, factory.produce, retire, factory.num_threads, factory.queue_size, factory.devices)
    return make_result(results) if make_result else results


def sweep(factory, labels=None, tags=None):
    """One-variable sweep of jobs returning `Series`: -> `Sweep(x = variable, y = series.variable, data)`
    (reference: pybinding/parallel.py:362-393)"""
    x = factory.variables[0]

    def make_result(data):
        first = data[0]
        if "energy" in factory.fixtures:      # the reference's convention: y is the factory's `energy` default (parallel.py:378-384)
            y = np.asarray(factory.fixtures["energy"])
        else:
            y = getattr(first, "variable", np.arange(np.size(getattr(first, "data", first))))
        rows = [np.asarray(getattr(d, "data", d)).squeeze() for d in data]
        return Sweep(x, y, np.vstack(rows), labels, tags)

    return parallel_for(factory, make_result)


def ndsweep(factory, labels=None, tags=None):
    """N-variable sweep: `NDSweep(variables + (energy,), data)` with `data.shape == [len(v) for v in variables]`
    (reference: pybinding/parallel.py:396-430, results.py:1027-1050).  Without an `energy` fixture the trailing axes are
    the shape of one result."""
    def make_result(data):
        arrays = [np.asarray(getattr(d, "data", d)).squeeze() for d in data]
        if "energy" in factory.fixtures:
            variables = list(factory.variables) + [np.asarray(factory.fixtures["energy"])]
            shape = tuple(len(v) for v in variables)
        else:
            variables = list(factory.variables)
            shape = tuple(len(v) for v in variables) + arrays[0].shape
        return NDSweep(variables, np.reshape(np.vstack([a.ravel() for a in arrays]), shape), labels, tags)

    return parallel_for(factory, make_result)
