import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


def random_ldu_mesh(rng, n, extra):
    """A random valid lduAddressing: a spanning chain plus `extra` random faces,
    unique (lower < upper) pairs in OpenFOAM's upper-triangular order."""
    pairs = {(i, i + 1) for i in range(n - 1)}
    while len(pairs) < n - 1 + extra:
        a, b = rng.integers(0, n, 2)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    arr = np.array(sorted(pairs), dtype=np.int32).reshape(-1, 2)
    return arr[:, 0].copy(), arr[:, 1].copy()


def gather_global(systems, xs):
    n = sum(s.n for s in systems)
    out = np.zeros(n)
    for s, x in zip(systems, xs):
        out[s.global_ids] = x
    return out
