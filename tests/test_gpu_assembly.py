"""-m gpu: device LDU -> CSR assembly is bit-exact against the oracle (and the
reference's golden vectors) through the C ABI."""
import os

import numpy as np
import pytest

from conftest import random_ldu_mesh
from gpu_helpers import upload_system
from ogl_b200 import cases, host
from ogl_b200.backend import Context, OglError

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context()
    yield c
    c.close()


def test_golden_vector_on_device(ctx):
    # unitTests/test_HostMatrix.C:70-107
    ctx.pattern_from_ldu(5, [0, 0, 1, 1, 2, 3], [1, 3, 2, 4, 3, 4], True)
    rows, cols, perm, rp = ctx.pattern_download()
    assert rows.tolist() == [0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4]
    assert cols.tolist() == [0, 1, 3, 0, 1, 2, 4, 1, 2, 3, 0, 2, 3, 4, 1, 3, 4]
    assert perm.tolist() == [6, 0, 1, 0, 7, 2, 3, 2, 8, 4, 1, 4, 9, 5, 3, 5, 10]
    assert rp.tolist() == [0, 3, 7, 10, 14, 17]
    # unitTests/test_HostMatrix.C:8-37 (values through the device gather)
    ctx.values_update([1., 2., 3., 4., 5.], [10., 11., 20., 12., 21., 13.])
    v, _ = ctx.values_download()
    # the golden permutation of that test differs from the one above; use ours
    stag = np.array([10., 11., 20., 12., 21., 13., 1., 2., 3., 4., 5.])
    assert np.array_equal(v, stag[perm])


def test_reference_fixtures_on_device(ctx):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ldu_ref_vectors.npz"))
    tags = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_rows")})
    for tag in tags:
        sym = tag.endswith("_sym")
        n = int(g[tag + "_n"])
        ctx.pattern_from_ldu(n, g[tag + "_lower"], g[tag + "_upper"], sym)
        rows, cols, perm, rp = ctx.pattern_download()
        assert np.array_equal(rows, g[tag + "_rows"])
        assert np.array_equal(cols, g[tag + "_cols"])
        assert np.array_equal(perm, g[tag + "_perm"])
        if float(g[tag + "_scale"]) == 1.0 or not sym:
            ctx.values_update(g[tag + "_diag"], g[tag + "_up"], None if sym else g[tag + "_lo"],
                              scaling=float(g[tag + "_scale"]))
            v, _ = ctx.values_download()
            assert np.array_equal(v, g[tag + "_vals"])


@pytest.mark.parametrize("n,extra", [(1, 0), (2, 0), (3, 1), (97, 300), (5000, 30000),
                                     (200000, 900000)])
def test_random_meshes_bit_exact(ctx, oracle, n, extra):
    rng = np.random.default_rng(n + extra)
    if n == 1:
        lower = upper = np.zeros(0, np.int32)
    else:
        lower, upper = random_ldu_mesh(rng, n, extra)
    F = lower.size
    for sym in (True, False):
        ctx.pattern_from_ldu(n, lower, upper, sym)
        rows, cols, perm, rp = ctx.pattern_download()
        o_rows, o_cols, o_perm = oracle.init_local_sparsity(n, upper, lower, sym)
        assert np.array_equal(rows, o_rows)
        assert np.array_equal(cols, o_cols)
        assert np.array_equal(perm, o_perm)
        assert np.array_equal(np.diff(rp), np.bincount(o_rows, minlength=n))
        diag, up, lo = rng.normal(size=n), rng.normal(size=F), rng.normal(size=F)
        for scale in (1.0, -0.5):
            ctx.values_update(diag, up, None if sym else lo, scaling=scale)
            v, _ = ctx.values_download()
            kind = "symmetric" if sym else "non_symmetric"
            assert np.array_equal(v, oracle.update_host(kind, o_perm, scale, diag, up, lo))


@pytest.mark.parametrize("builder", [
    lambda: cases.pressure_3d(24), lambda: cases.momentum_3d(20),
    lambda: cases.channel((16, 8, 8), (1, 1, 1)), lambda: cases.channel((16, 8, 8), (2, 2, 1)),
    lambda: cases.cavity_2d(), lambda: cases.pressure_3d(12, (2, 2, 2)),
    lambda: cases.build_case(cases.PressureModel((2, 3, 1), cyclic=(True, False, False)), (2, 3, 1)),
])
def test_cases_bit_exact_incl_interfaces(oracle, builder):
    for s in builder():
        c = Context(rank=0, n_ranks=1)
        try:
            ir, ic = host.collect_local_interface_indices(s)
            c.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, s.symmetric, ir, ic)
            # n_ranks = 1 context: exercise the halo pattern kernels without NCCL
            c.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
            c.values_update(s.diag, s.upper, None if s.symmetric else s.lower,
                            host.collect_interface_coeffs(s, True),
                            host.collect_interface_coeffs(s, False), 1.0)
            a = oracle.assemble(s)
            rows, cols, perm, rp = c.pattern_download()
            assert np.array_equal(rows, a.rows)
            assert np.array_equal(cols, a.cols)
            assert np.array_equal(perm, a.ldu_mapping)
            assert np.array_equal(rp, a.row_ptrs)
            v, nlv = c.values_download()
            assert np.array_equal(v, a.vals)
            if a.nl_rows.size:
                r, cc, m = c.nonlocal_pattern_download()
                assert np.array_equal(r, a.nl_rows)
                assert np.array_equal(cc, a.nl_cols)
                assert np.array_equal(m, a.nl_mapping)
                assert np.array_equal(nlv, a.nl_vals)
            tid, tsz, sidx = host.create_communication_pattern(s)
            assert np.array_equal(tid, a.target_ids)
            assert np.array_equal(tsz, a.target_sizes)
            assert np.array_equal(sidx, a.send_idxs)
        finally:
            c.close()


def test_bad_addressing_is_rejected(ctx):
    with pytest.raises(OglError):
        ctx.pattern_from_ldu(4, [0, 2], [1, 1], True)      # lower >= upper
    with pytest.raises(OglError):
        ctx.pattern_from_ldu(4, [0, 1], [1, 4], True)      # upper out of range
    with pytest.raises(OglError):
        ctx.values_update([1.0], [1.0])                    # no valid pattern any more


def test_structure_is_cached_only_values_move(ctx, oracle):
    s = cases.pressure_3d(16)[0]
    upload_system(ctx, s, partition=False)
    rows0, cols0, perm0, _ = ctx.pattern_download()
    s2 = cases.pressure_3d(16, sign=-1.0)[0]
    ctx.values_update(s2.diag, s2.upper)                   # second "time step"
    rows1, cols1, perm1, _ = ctx.pattern_download()
    assert np.array_equal(rows0, rows1) and np.array_equal(perm0, perm1)
    v, _ = ctx.values_download()
    assert np.array_equal(v, oracle.assemble(s2).vals)


def test_mtx_export_roundtrip(ctx, tmp_path):
    import scipy.io
    s = cases.momentum_3d(6)[0]
    upload_system(ctx, s, partition=False)
    p = tmp_path / "U_A_local.mtx"
    ctx.export_mtx(0, p)
    ctx.export_mtx(2, tmp_path / "U_rhs_b_.mtx")
    A = scipy.io.mmread(str(p)).tocsr()
    Aref, b = cases.assemble_global_csr([s])
    assert abs(A - Aref).max() <= 1e-14 * abs(Aref).max()
    bb = scipy.io.mmread(str(tmp_path / "U_rhs_b_.mtx")).ravel()
    assert np.allclose(bb, b, rtol=1e-14)
