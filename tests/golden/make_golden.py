"""Copies the reference's own KPM golden curves into a dependency-free .npz fixture.

Run in the build container (where /root/reference is mounted); the GPU box never reads /root/reference.
Source: /root/reference/tests/baseline_data/kpm/*.pbz = gzip + pickle (protocol 4) of float32 numpy arrays,
written by the reference's tests/test_kpm.py (test_ldos :25-47, test_dos :146-165, test_conductivity :176-197)
through tests/conftest.py:44-61 / pybinding/support/pickle.py:40-75.
"""
import glob
import gzip
import os
import pickle

import numpy as np

SRC = "/root/reference/tests/baseline_data/kpm"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kpm_baselines.npz")
# tests/test_parallel.py:16-52 -> baseline_data/parallel/{sweep,ndsweep}.pbz: pickled pybinding.results.Sweep / NDSweep
# objects (plain attribute dicts of numpy arrays); read without pybinding through a stand-in class
SRC_PARALLEL = "/root/reference/tests/baseline_data/parallel"
OUT_PARALLEL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_parallel_baselines.npz")


class _Stub:
    """Receives the attribute dict of any `pybinding.*` result class"""

    def __setstate__(self, state):   # pybinding/support/pickle.py:79-102: {"version": v, "dict": attributes}
        self.__dict__.update(state["dict"] if "version" in state else state)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("pybinding"):
            return _Stub
        return super().find_class(module, name)


if __name__ == "__main__":
    arrays = {}
    for path in sorted(glob.glob(os.path.join(SRC, "*.pbz"))):
        with gzip.open(path, "rb") as f:
            arrays[os.path.basename(path)[:-4]] = np.asarray(pickle.load(f))
    np.savez(OUT, **arrays)
    for k, v in arrays.items():
        print(k, v.dtype, v.shape)

    parallel = {}
    with gzip.open(os.path.join(SRC_PARALLEL, "sweep.pbz"), "rb") as f:
        sweep = _Unpickler(f).load()
    parallel.update({"sweep.x": np.asarray(sweep.x), "sweep.y": np.asarray(sweep.y), "sweep.data": np.asarray(sweep.data)})
    with gzip.open(os.path.join(SRC_PARALLEL, "ndsweep.pbz"), "rb") as f:
        nd = _Unpickler(f).load()
    for i, v in enumerate(nd.variables):
        parallel["ndsweep.variables.{}".format(i)] = np.asarray(v)
    parallel["ndsweep.data"] = np.asarray(nd.data)
    np.savez(OUT_PARALLEL, **parallel)
    for k, v in parallel.items():
        print(k, v.dtype, v.shape)
