"""Matrix-Market import: the files `ogl_export_mtx` / the
reference's `export` keyword write, read back into an LduSystem."""
import os

import numpy as np
import pytest

from ogl_b200 import cases, mtxio
from ogl_b200.host import FatalError


def dump(tmp_path, s, a, field="p"):
    """Write what the library exports, from the oracle's assembly of s."""
    mtxio.write_mtx_coordinate(tmp_path / f"{field}_A_local.mtx", a.n, a.n, a.rows, a.cols, a.vals)
    mtxio.write_mtx_coordinate(tmp_path / f"{field}_A_non_local.mtx", a.n, 0, [], [], [])
    mtxio.write_mtx_array(tmp_path / f"{field}_rhs_b_.mtx", s.source)


@pytest.mark.parametrize("builder", [lambda: cases.pressure_3d(7)[0], lambda: cases.momentum_3d(6)[0],
                                     lambda: cases.cavity_2d((1, 1, 1))[0]])
def test_round_trip_through_the_files(oracle, tmp_path, builder):
    s = builder()
    a = oracle.assemble(s)
    dump(tmp_path, s, a)
    t = mtxio.import_system(str(tmp_path), "p")
    assert t.n == s.n and t.symmetric == s.symmetric
    assert np.array_equal(t.lower_addr, s.lower_addr) and np.array_equal(t.upper_addr, s.upper_addr)
    # 15 significant digits in the files
    assert np.allclose(t.diag, s.diag, rtol=1e-14, atol=0) and np.allclose(t.upper, s.upper, rtol=1e-14, atol=0)
    if not s.symmetric:
        assert np.allclose(t.lower, s.lower, rtol=1e-14, atol=0)
    assert np.allclose(t.source, s.source, rtol=1e-14, atol=1e-300)
    # and the re-assembled matrix has the pattern and permutation of the original
    b = oracle.assemble(t)
    assert np.array_equal(a.rows, b.rows) and np.array_equal(a.cols, b.cols)
    assert np.array_equal(a.ldu_mapping, b.ldu_mapping)


def test_cyclic_entries_become_internal_faces(oracle, tmp_path):
    # the channel's cyclic couplings are ordinary entries of A_local: the imported system has
    # no interfaces but assembles to the same matrix
    s = cases.channel((8, 4, 4), (1, 1, 1))[0]
    a = oracle.assemble(s)
    dump(tmp_path, s, a)
    t = mtxio.import_system(str(tmp_path), "p")
    b = oracle.assemble(t)
    assert np.array_equal(a.rows, b.rows) and np.array_equal(a.cols, b.cols)
    assert np.allclose(a.vals, b.vals, rtol=1e-14, atol=0)


def test_rejects_what_is_not_an_ldu_matrix(tmp_path):
    mtxio.write_mtx_coordinate(tmp_path / "p_A_local.mtx", 3, 3, [0, 1, 2, 0], [0, 1, 2, 2], [1, 1, 1, 5])
    with pytest.raises(FatalError, match="structurally symmetric"):
        mtxio.import_system(str(tmp_path), "p")
    mtxio.write_mtx_coordinate(tmp_path / "p_A_local.mtx", 3, 3, [0, 1, 0, 1], [0, 1, 1, 0], [1, 1, 2, 2])
    with pytest.raises(FatalError, match="diagonal"):
        mtxio.import_system(str(tmp_path), "p")
    mtxio.write_mtx_coordinate(tmp_path / "p_A_local.mtx", 2, 2, [0, 1], [0, 1], [1, 1])
    mtxio.write_mtx_coordinate(tmp_path / "p_A_non_local.mtx", 2, 1, [0], [0], [3])
    with pytest.raises(FatalError, match="non-local"):
        mtxio.import_system(str(tmp_path), "p")
    (tmp_path / "q_A_local.mtx").write_text("not a matrix\n")
    with pytest.raises(FatalError, match="Matrix-Market"):
        mtxio.import_system(str(tmp_path), "q")


@pytest.mark.gpu
def test_export_import_solve_again(oracle, tmp_path):
    """Device export -> import -> solve: same iteration count, solution within the file precision."""
    from ogl_b200.host import ObjectRegistry
    from ogl_b200.parallel import Pstream
    from ogl_b200.plugin import lduMatrix_solver_New

    s = cases.pressure_3d(16)[0]
    controls = {"solver": "GKOCG", "executor": "cuda", "tolerance": 1e-9, "relTol": 0.0,
                "adaptMinIter": False, "debug": True, "writeTime": True, "preconditioner": "BJ"}
    db = ObjectRegistry()
    db["__time_path__"] = str(tmp_path / "processor0" / "0.005")
    sol = lduMatrix_solver_New("p", s, controls, db, Pstream())
    psi = s.psi.copy()
    perf = sol.solve(psi, s.source)
    folder = None
    for root, _, files in os.walk(tmp_path):
        if "p_A_local.mtx" in files:
            folder = root
    assert folder is not None, "export keyword wrote no files"
    t = mtxio.import_system(folder, "p")
    controls2 = dict(controls, debug=False)
    sol2 = lduMatrix_solver_New("p2", t, controls2, ObjectRegistry(), Pstream())
    psi2 = t.psi.copy()
    perf2 = sol2.solve(psi2, t.source)
    assert abs(perf2.n_iterations - perf.n_iterations) <= 2
    assert np.linalg.norm(psi2 - psi) / np.linalg.norm(psi) <= 1e-8


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_decomposed_dump_round_trip(oracle, tmp_path, procs):
    """Per-processor dumps + the partition side-car -> the ranks' systems with their processor
    interfaces: same assembled matrices, same communication pattern, same solve."""
    from ogl_b200 import host
    systems = cases.pressure_3d(8, procs)
    asms = [oracle.assemble(s) for s in systems]
    for s, a in zip(systems, asms):
        folder = tmp_path / f"processor{s.rank}" / "0.005"
        os.makedirs(folder)
        mtxio.write_mtx_coordinate(folder / "p_A_local.mtx", a.n, a.n, a.rows, a.cols, a.vals)
        mtxio.write_mtx_coordinate(folder / "p_A_non_local.mtx", a.n, a.nl_rows.size, a.nl_rows, a.nl_cols, a.nl_vals)
        mtxio.write_mtx_array(folder / "p_rhs_b_.mtx", s.source)
        tid, tsz, _ = host.create_communication_pattern(s)
        mtxio.write_partition_sidecar(str(folder), "p", s.rank, len(systems), s.n, tid, tsz)
    back = mtxio.import_decomposed(str(tmp_path), "0.005", "p")
    assert len(back) == len(systems)
    basms = [oracle.assemble(t) for t in back]
    for a, b in zip(asms, basms):
        assert np.array_equal(a.rows, b.rows) and np.array_equal(a.cols, b.cols)
        assert np.allclose(a.vals, b.vals, rtol=1e-14, atol=0)
        assert np.array_equal(a.target_ids, b.target_ids) and np.array_equal(a.target_sizes, b.target_sizes)
        assert np.array_equal(a.send_idxs, b.send_idxs)
        assert np.array_equal(a.nl_rows, b.nl_rows) and np.allclose(a.nl_vals, b.nl_vals, rtol=1e-14, atol=0)
    o1 = oracle.solve(asms, "GKOCG", "BJ", tolerance=1e-9)
    o2 = oracle.solve(basms, "GKOCG", "BJ", tolerance=1e-9)
    assert abs(o1.n_iterations - o2.n_iterations) <= 1
    for x1, x2 in zip(o1.x, o2.x):
        assert np.allclose(x1, x2, rtol=0, atol=1e-9 * np.abs(x1).max())


def test_decomposed_dump_without_sidecar_is_rejected(tmp_path):
    folder = tmp_path / "processor0" / "1"
    os.makedirs(folder)
    mtxio.write_mtx_coordinate(folder / "p_A_local.mtx", 2, 2, [0, 1], [0, 1], [1, 1])
    mtxio.write_mtx_coordinate(folder / "p_A_non_local.mtx", 2, 1, [0], [0], [3])
    with pytest.raises(FatalError, match="side-car"):
        mtxio.import_decomposed(str(tmp_path), "1", "p")
