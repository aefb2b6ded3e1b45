"""The oracle's assembly half pinned against (1) the reference's own
known-answer tests, (2) fixtures produced by the reference's free functions
compiled from source, (3) that compiled reference itself when present."""
import os

import numpy as np
import pytest

from conftest import random_ldu_mesh

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ldu_ref_vectors.npz")


# unitTests/test_HostMatrix.C:70-107
def test_init_local_sparsity_known_answer(oracle):
    rows, cols, perm = oracle.init_local_sparsity(5, [1, 3, 2, 4, 3, 4], [0, 0, 1, 1, 2, 3], True)
    assert rows.tolist() == [0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4]
    assert cols.tolist() == [0, 1, 3, 0, 1, 2, 4, 1, 2, 3, 0, 2, 3, 4, 1, 3, 4]
    assert perm.tolist() == [6, 0, 1, 0, 7, 2, 3, 2, 8, 4, 1, 4, 9, 5, 3, 5, 10]


# unitTests/test_HostMatrix.C:8-37
def test_symmetric_update_known_answer(oracle):
    d = [1., 2., 3., 4., 5.]
    u = [10., 11., 20., 12., 21., 13.]
    p = [6, 0, 2, 0, 7, 1, 4, 1, 8, 3, 2, 3, 9, 5, 4, 5, 10]
    exp = [1., 10., 20., 10., 2., 11., 21., 11., 3., 12., 20., 12., 4., 13., 21., 13., 5.]
    for which in ("oracle", "as_written"):
        assert oracle.update_host("symmetric", p, 1.0, d, u, which=which).tolist() == exp


# unitTests/test_HostMatrix.C:39-68
def test_non_symmetric_update_known_answer(oracle):
    d = [1.] * 5
    u = [1., 2., 1., 2., 1., 1.]
    l = [2., 2., 3., 2., 3., 2.]
    p = [12, 0, 1, 6, 13, 2, 3, 7, 14, 4, 8, 9, 15, 5, 10, 11, 16]
    exp = [1., 1., 2., 2., 1., 1., 2., 2., 1., 1., 3., 2., 1., 1., 3., 2., 1.]
    assert oracle.update_host("non_symmetric", p, 1.0, d, u, l).tolist() == exp


def test_fixtures_from_compiled_reference(oracle):
    g = np.load(GOLD)
    tags = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_rows")})
    assert len(tags) == 10
    for tag in tags:
        sym = tag.endswith("_sym")
        n = int(g[tag + "_n"])
        rows, cols, perm = oracle.init_local_sparsity(n, g[tag + "_upper"], g[tag + "_lower"], sym)
        assert np.array_equal(rows, g[tag + "_rows"])
        assert np.array_equal(cols, g[tag + "_cols"])
        assert np.array_equal(perm, g[tag + "_perm"])
        scale = float(g[tag + "_scale"])
        if sym:
            # the reference's symmetric_update drops `scale` (operator precedence,
            # HostMatrixFreeFunctions.C:27-28): the as-written twin must match it
            v = oracle.update_host("symmetric", perm, scale, g[tag + "_diag"], g[tag + "_up"],
                                   which="as_written")
        else:
            v = oracle.update_host("non_symmetric", perm, scale, g[tag + "_diag"], g[tag + "_up"],
                                   g[tag + "_lo"])
        assert np.array_equal(v, g[tag + "_vals"])


def test_oracle_equals_compiled_reference_on_random_meshes(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libogl_ref.so not built (reference tree absent)")
    rng = np.random.default_rng(7)
    for n, extra in [(2, 0), (3, 1), (50, 40), (333, 1200), (5000, 30000)]:
        lower, upper = random_ldu_mesh(rng, n, extra)
        F = lower.size
        for sym in (True, False):
            a = oracle.init_local_sparsity(n, upper, lower, sym)
            b = oracle.init_local_sparsity(n, upper, lower, sym, which="ref")
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
            perm = a[2]
            diag, up, lo = rng.normal(size=n), rng.normal(size=F), rng.normal(size=F)
            n_if = 7
            iface = rng.normal(size=n_if)
            # permutation with an interface section appended (w_interface variants)
            base = F if sym else 2 * F
            perm_if = np.concatenate([perm, base + n + np.arange(n_if)]).astype(np.int32)
            for scale in (1.0, -1.0, 0.37):
                if sym:
                    assert np.array_equal(
                        oracle.update_host("symmetric", perm, scale, diag, up, which="as_written"),
                        oracle.update_host("symmetric", perm, scale, diag, up, which="ref"))
                    assert np.array_equal(
                        oracle.update_host("symmetric_w_interface", perm_if, scale, diag, up,
                                           iface=iface),
                        oracle.update_host("symmetric_w_interface", perm_if, scale, diag, up,
                                           iface=iface, which="ref"))
                else:
                    assert np.array_equal(
                        oracle.update_host("non_symmetric", perm, scale, diag, up, lo),
                        oracle.update_host("non_symmetric", perm, scale, diag, up, lo, which="ref"))
                    assert np.array_equal(
                        oracle.update_host("non_symmetric_w_interface", perm_if, scale, diag, up,
                                           lo, iface),
                        oracle.update_host("non_symmetric_w_interface", perm_if, scale, diag, up,
                                           lo, iface, which="ref"))


def test_scaled_symmetric_update_is_the_documented_intent(oracle):
    # README.md:81 (sAx = sb): with scale != 1 the intent differs from the as-written code
    rows, cols, perm = oracle.init_local_sparsity(5, [1, 3, 2, 4, 3, 4], [0, 0, 1, 1, 2, 3], True)
    d = np.arange(1., 6.)
    u = np.array([10., 11., 20., 12., 21., 13.])
    v1 = oracle.update_host("symmetric", perm, 1.0, d, u)
    v2 = oracle.update_host("symmetric", perm, -2.0, d, u)
    assert np.array_equal(v2, -2.0 * v1)


def test_face_less_mesh(oracle):
    # n = 1, F = 0: undefined in the reference (HostMatrixFreeFunctions.C:157-158)
    rows, cols, perm = oracle.init_local_sparsity(1, [], [], True)
    assert rows.tolist() == [0] and cols.tolist() == [0] and perm.tolist() == [0]


def test_assemble_matches_scipy_structure(oracle):
    import scipy.sparse as sp
    from ogl_b200 import cases

    for systems in (cases.pressure_3d(6), cases.momentum_3d(5), cases.channel((8, 4, 4), (1, 1, 1))):
        s = systems[0]
        a = oracle.assemble(s)
        # row-major, strictly increasing (row, col) except cyclic duplicates
        key = a.rows.astype(np.int64) * s.n + a.cols
        assert np.all(np.diff(key) >= 0)
        A = sp.coo_matrix((a.vals, (a.rows, a.cols)), shape=(s.n, s.n)).tocsr()
        Aref, b = cases.assemble_global_csr(systems)
        assert abs(A - Aref).max() == 0.0
        assert np.array_equal(b, a.b)


def test_cyclic_merge_puts_interface_after_equal_entry(oracle):
    # 2 cells wide in the cyclic direction: the cyclic coupling duplicates an
    # existing (row, col); HostMatrix.C:543-575 inserts it AFTER the existing entry
    from ogl_b200 import cases

    m = cases.PressureModel((2, 3, 1), cyclic=(True, False, False))
    s = cases.build_case(m, (2, 3, 1))[0]
    a = oracle.assemble(s)
    F, n = s.n_faces, s.n
    key = a.rows.astype(np.int64) * n + a.cols
    dup = np.nonzero(np.diff(key) == 0)[0]
    assert dup.size > 0
    for k in dup:
        assert a.ldu_mapping[k] < F + n <= a.ldu_mapping[k + 1]


def test_comm_and_non_local_pattern(oracle):
    from ogl_b200 import cases

    systems = cases.channel((8, 4, 4), (2, 2, 1))
    asms = [oracle.assemble(s) for s in systems]
    for s, a in zip(systems, asms):
        proc = [i for i in s.interfaces if i.kind == "processor"]
        assert np.all(np.diff(a.target_ids) > 0)
        assert a.target_sizes.sum() == a.nl_rows.size == sum(p.face_cells.size for p in proc)
        # rows ascending, ties in running-index order (stable)
        assert np.all(np.diff(a.nl_rows) >= 0)
        same = np.diff(a.nl_rows) == 0
        assert np.all(np.diff(a.nl_cols)[same] > 0)
        assert np.array_equal(a.nl_cols, a.nl_mapping)
        fcs = np.concatenate([p.face_cells for p in proc])
        assert np.array_equal(fcs[a.nl_cols], a.nl_rows)
    # what rank r sends to q is what q expects from r, block by block
    for r, a in enumerate(asms):
        off = 0
        for t, q in enumerate(a.target_ids):
            aq = asms[q]
            u = list(aq.target_ids).index(r)
            assert aq.target_sizes[u] == a.target_sizes[t]
            off += a.target_sizes[t]


def test_preconditioner_restatements_reproduce_their_regression_fixtures(oracle):
    """tests/golden/precond_oracle_vectors.npz (made by tests/golden/make_precond_golden.py) pins the
    oracle's ILU / IC factors and applies, the IRILU apply and the multigrid aggregates / first coarse
    matrix / V cycle on tiny systems against accidental change.  A regression fixture of the oracle itself:
    these restatements have no reference vector to be pinned against (Ginkgo is absent)."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_precond_golden", os.path.join(path, "make_precond_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    gold = np.load(os.path.join(path, "precond_oracle_vectors.npz"))
    assert set(gold.files) == set(now)
    for key in gold.files:
        assert np.array_equal(gold[key], now[key]), key
