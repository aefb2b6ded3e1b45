"""CPU: host-side logic of the plugin layer (no device calls): communication
pattern, interface collection, keyword defaults, adaptive criterion, the
per-field key-value store, run-time selection errors, case generators."""
import numpy as np
import pytest

from ogl_b200 import cases, host, plugin
from ogl_b200.host import ObjectRegistry


def test_communication_pattern_matches_oracle(oracle):
    for systems in (cases.channel((8, 4, 4), (2, 2, 1)), cases.pressure_3d(6, (2, 2, 2)),
                    cases.cavity_2d()):
        for s in systems:
            tid, tsz, sidx = host.create_communication_pattern(s)
            a = oracle.assemble(s)
            assert np.array_equal(tid, a.target_ids)
            assert np.array_equal(tsz, a.target_sizes)
            assert np.array_equal(sidx, a.send_idxs)
            assert host.count_interface_nnz(s, True) == a.nl_rows.size
            assert host.count_interface_nnz(s, False) == a.rows.size - s.n - 2 * s.n_faces


def test_two_patches_to_the_same_neighbour_are_concatenated():
    # cyclic in x cut by a [2,1,1] decomposition: plain + processorCyclic patch
    s0, s1 = cases.channel((8, 4, 4), (2, 1, 1))
    proc0 = [i for i in s0.interfaces if i.kind == "processor"]
    assert [p.nbr_rank for p in proc0] == [1, 1]
    tid, tsz, sidx = host.create_communication_pattern(s0)
    assert tid.tolist() == [1] and tsz.tolist() == [2 * 16]
    assert np.array_equal(sidx, np.concatenate([p.face_cells for p in proc0]))


def test_local_interface_indices():
    s = cases.channel((8, 4, 4), (1, 1, 1))[0]
    rows, cols = host.collect_local_interface_indices(s)
    cyc = [i for i in s.interfaces if i.kind == "cyclic"]
    assert len(cyc) == 4 and rows.size == sum(i.face_cells.size for i in cyc)
    # a cyclic pair maps onto each other
    for i in cyc:
        j = s.interfaces[i.nbr_patch]
        assert s.interfaces[j.nbr_patch] is i
    s.interfaces[-1].kind = "cyclicAMI"
    with pytest.raises(host.FatalError):
        host.collect_local_interface_indices(s)


def test_stopping_criterion_defaults_and_adaptation():
    c = plugin.StoppingCriterion({"solver": "GKOCG"})
    # StoppingCriterion.H:164-177 (relTol default is 1e-6, relaxationFactor 0.6)
    assert (c.max_iter, c.min_iter, c.tolerance, c.rel_tol) == (1000, 0, 1e-6, 1e-6)
    assert (c.frequency, c.relaxation, c.adapt_min_iter, c.norm_eval_limit) == (1, 0.6, True, 100)
    assert plugin.StoppingCriterion({"solver": "GKOBiCGStab"}).max_iter == 2000   # :188
    # :199-209
    assert c.effective(False, 0, 5.0) == (0, 1)
    assert c.effective(True, 100, 5.0) == (0, 1)          # export disables adaptation
    mi, fr = c.effective(False, 100, 4.0)
    assert mi == 60
    alpha = (1.0 / (100 * 0.4) * 4.0) ** 0.5
    assert fr == min(100, max(1, int(1 / alpha)))
    assert c.effective(False, 100000, 1.0)[1] == 100      # capped by normEvalLimit
    off = plugin.StoppingCriterion({"solver": "GKOCG", "adaptMinIter": False})
    assert off.effective(False, 100, 4.0) == (0, 1)


def test_key_value_store_truncates_like_the_reference():
    db = ObjectRegistry()
    # common.C:97,145: default 1 for the iteration count, label-typed storage
    assert plugin.get_solve_prev_iters("p", db, True) == 1
    plugin.set_solve_prev_iters("p", db, 57, True)
    assert plugin.get_solve_prev_iters("p", db, True) == 57
    assert plugin.get_solve_prev_iters("p", db, False) == 1
    plugin.set_solve_prev_rel_res_cost("p", db, 3.9)
    assert plugin.get_solve_prev_rel_res_cost("p", db) == 3       # truncated (Appendix B-6)
    plugin.set_solve_prev_rel_res_cost("p", db, 0.7)
    assert plugin.get_solve_prev_rel_res_cost("p", db) == 0       # adaptation off


def test_runtime_selection_errors_before_any_device_call():
    db = ObjectRegistry()
    sym = cases.pressure_3d(4)[0]
    asym = cases.momentum_3d(4)[0]
    base = {"preconditioner": "BJ", "executor": "cuda"}
    with pytest.raises(host.FatalError, match="Unknown asymmetric matrix solver GKOCG"):
        plugin.lduMatrix_solver_New("U", asym, dict(base, solver="GKOCG"), db)
    with pytest.raises(host.FatalError, match="Unknown symmetric matrix solver PCG"):
        plugin.lduMatrix_solver_New("p", sym, dict(base, solver="PCG"), db)
    with pytest.raises(host.FatalError, match="does not support the executor"):
        plugin.lduMatrix_solver_New("p", sym, dict(base, solver="GKOCG", executor="reference"), db)


def test_case_generators():
    import scipy.sparse as sp
    # hex-mesh counts of SURVEY section 8: F = 3 N^2 (N-1)
    lo, up, d = cases.box_addressing(10, 10, 10)
    assert lo.size == 3 * 100 * 9 and np.all(lo < up)
    key = lo.astype(np.int64) * 1000 + up
    assert np.all(np.diff(key) > 0)              # upper-triangular order
    # decomposed == undecomposed global operator
    for procs in [(2, 1, 1), (2, 2, 1), (2, 2, 2)]:
        A1, b1 = cases.assemble_global_csr(cases.pressure_3d(6))
        An, bn = cases.assemble_global_csr(cases.pressure_3d(6, procs))
        # rhs: same terms, summed in a different order on the cut rows
        assert abs(A1 - An).max() == 0 and np.allclose(b1, bn, rtol=1e-13, atol=1e-20)
    A, b = cases.assemble_global_csr(cases.momentum_3d(6, (2, 2, 1)))
    A1, b1 = cases.assemble_global_csr(cases.momentum_3d(6))
    # the diagonal accumulates its face terms in a different order on cut rows
    assert abs(A - A1).max() <= 1e-14 * abs(A1).max() and np.allclose(b, b1, rtol=1e-12)
    d = A.diagonal()
    off = abs(A).sum(axis=1).A1 - abs(d)
    assert np.all(d > off)                       # strictly diagonally dominant
    # cavity: sign/magnitude bounds of test/data_validation.py:93-111 on the SPD twin
    for s in cases.cavity_2d(sign=-1.0):
        assert np.all((s.diag >= 0) & (s.diag <= 1e-3))
        assert np.all((s.upper <= 0) & (s.upper >= -1e-3))
    # cyclic channel is symmetric and couples x=0 with x=Nx-1
    A, _ = cases.assemble_global_csr(cases.channel((8, 4, 4), (1, 1, 1)))
    assert abs(A - A.T).max() == 0 and A[0, 7] != 0
