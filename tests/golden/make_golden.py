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

if __name__ == "__main__":
    arrays = {}
    for path in sorted(glob.glob(os.path.join(SRC, "*.pbz"))):
        with gzip.open(path, "rb") as f:
            arrays[os.path.basename(path)[:-4]] = np.asarray(pickle.load(f))
    np.savez(OUT, **arrays)
    for k, v in arrays.items():
        print(k, v.dtype, v.shape)
