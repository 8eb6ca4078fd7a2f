import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import numpy as np
    import scipy.sparse as sp

    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    for k in ("A", "B"):
        d[k] = sp.csr_matrix((d[k + "_data"], d[k + "_indices"], d[k + "_indptr"]), shape=tuple(d[k + "_shape"]))
    return d


def corr_from_array(arr):
    data = {}
    for i, j, xi, eta in arr:
        data.setdefault(int(i), []).append((int(j), float(xi), float(eta)))
    return data


def align_signs(Phi, ref):
    import numpy as np

    s = np.sign(np.einsum("ij,ij->j", Phi, ref))
    s[s == 0] = 1.0
    return Phi * s, s
