"""-m gpu tests of the merge-path SpMV (spmv_variant 8, ogl_b200/csrc/spmv_merge.cu) and of the long-row path
of the device assembly that were WRITTEN AFTER THIS ROUND'S GPU BUDGET WAS SPENT: they have not run on
hardware yet, so they live outside tests/ (the round-end `pytest tests -m gpu` stays the verified set).
What did run on a B200: tests/test_gpu_spmv.py::test_spmv_one_long_row_takes_the_merge_path_kernel (automatic
selection, full and ragged slices, the thread-per-row branch, one row split over two slices and its carry).

    python -m pytest tools/pending_gpu_tests -m gpu -q        # on a B200, from the repo root
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import upload_system  # noqa: E402
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context, OglError  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="module")
def ctx():
    c = Context()
    yield c
    c.close()


def _arrow_system(n, seed=2):
    """Cell 0 shares a face with every other cell, the others form a chain: one row of n entries among
    rows of 4 (SPD M-matrix)."""
    lower = np.concatenate([np.zeros(n - 1, np.int32), np.arange(1, n - 1, dtype=np.int32)])
    upper = np.concatenate([np.arange(1, n, dtype=np.int32), np.arange(2, n, dtype=np.int32)])
    rng = np.random.default_rng(seed)
    up = -rng.uniform(0.5, 1.0, lower.size)
    diag = np.full(n, 0.05)
    np.add.at(diag, lower, -up)
    np.add.at(diag, upper, -up)
    x_star = rng.uniform(-1, 1, n)
    b = diag * x_star
    np.add.at(b, lower, up * x_star[upper])
    np.add.at(b, upper, up * x_star[lower])
    return cases.LduSystem(n=n, lower_addr=lower, upper_addr=upper, diag=diag, upper=up, lower=None, interfaces=[],
                           source=b, psi=np.zeros(n), global_ids=np.arange(n, dtype=np.int64), x_star=x_star)


def test_merge_path_kernel_on_an_irregular_matrix(ctx, oracle):
    """spmv_variant 8 (spmv_merge.cu), picked from the row statistics: one row spans ~6 slices of 2048
    entries, its carries are added in slice order; every other row is summed left to right.  SpMV to
    1e-13 (the split row's association differs), the same result run to run, and a CG solve whose
    prologue (advanced apply) and fused <p,q> go through the same kernel."""
    from gpu_helpers import gpu_solve, rel_l2
    s = _arrow_system(12000)
    upload_system(ctx, s, partition=False)
    assert ctx.get_option("spmv_variant_in_use") == 8
    a = oracle.assemble(s)
    x = np.random.default_rng(3).normal(size=s.n)
    y, y_ref = ctx.spmv(x), oracle.dist_spmv([a], [x])[0]
    assert np.allclose(y, y_ref, rtol=1e-13, atol=1e-13 * np.abs(y_ref).max())
    assert (y != y_ref).sum() <= ctx.nnz // 2048 + 2   # only rows cut by a slice boundary may differ at all
    assert np.array_equal(ctx.spmv(x), y)
    r, xs = gpu_solve(ctx, "GKOCG", "BJ", tolerance=1e-9)
    o = oracle.solve([a], "GKOCG", "BJ", tolerance=1e-9)
    assert abs(r.n_iterations - o.n_iterations) <= 2 and rel_l2(xs, o.x[0]) <= 1e-8


@pytest.mark.parametrize("builder,solver", [(lambda: cases.pressure_3d(40)[0], "GKOCG"),
                                            (lambda: cases.momentum_3d(24)[0], "GKOBiCGStab"),
                                            (lambda: cases.channel((16, 8, 8), (1, 1, 1))[0], "GKOGMRES")])
def test_merge_path_kernel_forced_on_mesh_matrices(ctx, oracle, builder, solver):
    """The same kernel forced onto regular matrices (sizes that are not a multiple of the slice length,
    rows split across slice boundaries): SpMV to 1e-13, solves within the north-star bars -- CG (one fused
    reduction), BiCGStab (two) and GMRES."""
    from gpu_helpers import gpu_solve, rel_l2
    s = builder()
    upload_system(ctx, s, partition=False)
    ctx.set_option("spmv_variant", 8)
    a = oracle.assemble(s)
    x = np.random.default_rng(4).normal(size=s.n)
    y, y_ref = ctx.spmv(x), oracle.dist_spmv([a], [x])[0]
    assert np.allclose(y, y_ref, rtol=1e-13, atol=1e-13 * np.abs(y_ref).max())
    assert (y != y_ref).mean() < 0.01                 # only rows cut by a slice boundary may differ
    kw = {"krylov_dim": 30} if solver == "GKOGMRES" else {}
    r, xs = gpu_solve(ctx, solver, "BJ", tolerance=1e-9, **kw)
    o = oracle.solve([a], solver, "BJ", tolerance=1e-9, **kw)
    assert abs(r.n_iterations - o.n_iterations) <= 2 and rel_l2(xs, o.x[0]) <= 1e-8
    ctx.set_option("spmv_variant", 0)


def test_long_rows_assemble_bit_exact(ctx, oracle):
    """Rows beyond 4096 entries leave the per-thread insertion sort of the device assembly for a radix sort
    of their (column, slot) keys (assembly.cu:sort_long_rows): rows / cols / ldu_mapping stay bit-identical
    to the oracle, and the merge-path SpMV over the result matches to 1e-13."""
    s = _arrow_system(300000)
    upload_system(ctx, s, partition=False)
    a = oracle.assemble(s)
    rows, cols, perm, _ = ctx.pattern_download()
    assert np.array_equal(rows, a.rows) and np.array_equal(cols, a.cols) and np.array_equal(perm, a.ldu_mapping)
    assert ctx.get_option("max_row_len") == s.n and ctx.get_option("spmv_variant_in_use") == 8
    x = np.random.default_rng(6).normal(size=s.n)
    y, y_ref = ctx.spmv(x), oracle.dist_spmv([a], [x])[0]
    assert np.allclose(y, y_ref, rtol=1e-13, atol=1e-13 * np.abs(y_ref).max())
