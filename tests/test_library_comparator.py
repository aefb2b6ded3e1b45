"""The stock-library composition of the PCG iteration (tools/cusparse_cg.py: torch.sparse_csr
SpMV + library BLAS-1, SURVEY section 8d "comparator") must describe the same iteration as the
oracle -- a third implementation written against a different substrate."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from ogl_b200 import cases

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("n", [12, 20])
def test_library_pcg_matches_the_oracle(oracle, n):
    import cusparse_cg

    s = cases.pressure_3d(n)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # torch: sparse CSR support is in beta
        A = cusparse_cg.csr_of(s, torch.device("cpu"))
        b = torch.from_numpy(s.source.copy())
        x = torch.zeros_like(b)
        calls, init, final, _ = cusparse_cg.pcg(A, b, x, 1.0 / torch.from_numpy(s.diag.copy()), tolerance=1e-7)
    o = oracle.solve([oracle.assemble(s)], "GKOCG", "BJ", tolerance=1e-7)
    assert abs(calls - o.criterion_calls) <= 1
    assert init == pytest.approx(o.init_residual, rel=1e-10)
    assert final == pytest.approx(o.final_residual, rel=1e-5) or calls != o.criterion_calls
    assert np.linalg.norm(x.numpy() - o.x[0]) <= 1e-8 * np.linalg.norm(o.x[0])
