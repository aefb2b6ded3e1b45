"""-m gpu: the `preconditioner Multigrid` keyword (Preconditioner.H:261-341) through
the C ABI against the oracle's restatement (oracle/multigrid.hpp).

Bars: the hierarchy -- aggregates, coarse sparsity patterns and coarse coefficients of every level --
is BIT-EXACT (integer work, and FP64 sums taken in the oracle's order); one V cycle agrees to 1e-12
(the coarsest CG's dot products are tree sums on the device); solves keep the north-star bars
(iterations +-2, relative L2 difference <= 1e-8)."""
import numpy as np
import pytest

from gpu_helpers import PRECOND_ID, gpu_solve, rel_l2, upload_system
from ogl_b200 import cases
from ogl_b200.backend import Context
from ogl_b200.host import FatalError, ObjectRegistry
from ogl_b200.plugin import lduMatrix_solver_New
from test_gpu_trifactor import SYSTEMS, _csr

pytestmark = pytest.mark.gpu

ITER_TOL = 2
L2_TOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    c = Context()
    yield c
    c.close()


@pytest.mark.parametrize("system", ["momentum", "pressure_spd", "channel", "unstructured"])
def test_hierarchy_bit_exact_and_cycle(ctx, oracle, system):
    s = SYSTEMS[system]()
    upload_system(ctx, s, partition=False)
    a, rp = _csr(oracle, s)
    H = oracle.MgHierarchy(s.n, rp, a.cols, a.vals)
    ctx.precond_setup(PRECOND_ID["Multigrid"], 1)
    levels = ctx.mg_levels()
    assert [(l["n"], l["nnz"], l["n_coarse"]) for l in levels] == [(l["n"], l["nnz"], l["n_coarse"]) for l in H.levels]
    for dev, ref in zip(levels, H.levels):
        assert np.array_equal(dev["row_ptrs"], ref["row_ptrs"])
        assert np.array_equal(dev["cols"], ref["cols"])
        assert np.array_equal(dev["vals"], ref["vals"])
        assert (dev["agg"] is None) == (ref["agg"] is None)
        if ref["agg"] is not None:
            assert np.array_equal(dev["agg"], ref["agg"])
    r = np.random.default_rng(21).standard_normal(s.n)
    z, z_ref = ctx.precond_apply(r), H.apply(r)
    assert np.linalg.norm(z - z_ref) <= 1e-12 * np.linalg.norm(z_ref)
    assert np.array_equal(ctx.precond_apply(r), z)          # run-to-run identical


@pytest.mark.parametrize("solver,system,kw", [
    ("GKOCG", "pressure_spd", {}),
    ("GKOBiCGStab", "momentum", {}),
    ("GKOGMRES", "momentum", {"krylov_dim": 30}),
    ("GKOCG", "unstructured", {}),
])
def test_solves_match_oracle(ctx, oracle, solver, system, kw):
    s = SYSTEMS[system]()
    upload_system(ctx, s, partition=False)
    r, x = gpu_solve(ctx, solver, "Multigrid", 1, tolerance=1e-9, **kw)
    o = oracle.solve([oracle.assemble(s)], solver, "Multigrid", tolerance=1e-9, **kw)
    assert abs(r.n_iterations - o.n_iterations) <= ITER_TOL, (r.n_iterations, o.n_iterations)
    assert rel_l2(x, o.x[0]) <= L2_TOL
    assert r.final_residual < 1e-9
    upload_system(ctx, s, partition=False)
    rj, _ = gpu_solve(ctx, solver, "BJ", tolerance=1e-9, **kw)
    assert r.n_iterations < rj.n_iterations


def test_graph_and_stream_paths_agree(ctx):
    s = cases.pressure_3d(20, sign=-1.0)[0]
    out = []
    for use_graph in (1, 0):
        ctx.set_option("use_graph", use_graph)
        upload_system(ctx, s, partition=False)
        r, x = gpu_solve(ctx, "GKOCG", "Multigrid", tolerance=1e-9)
        out.append((r.n_iterations, x))
    ctx.set_option("use_graph", 1)
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


def test_plugin_keywords(oracle):
    """Sub-dictionary keywords reach the hierarchy (maxLevels, minCoarseRows, coarseSolverIters); with
    `caching` a later solve reuses the hierarchy; cycle w / f are rejected."""
    s = cases.pressure_3d(16)[0]
    spd = cases.pressure_3d(16, sign=-1.0)[0]
    pre = {"preconditioner": "Multigrid", "maxLevels": 3, "minCoarseRows": 50, "coarseSolverIters": 8, "caching": 1}
    o = oracle.solve([oracle.assemble(spd)], "GKOCG", "Multigrid", tolerance=1e-9, mg_max_levels=3,
                     mg_min_coarse_rows=50, mg_coarse_iters=8)
    controls = {"solver": "GKOCG", "executor": "cuda", "tolerance": 1e-9, "relTol": 0.0, "adaptMinIter": False,
                "scaling": -1.0, "updateInitGuess": True, "preconditioner": pre}
    db = ObjectRegistry()
    sol = lduMatrix_solver_New("p", s, controls, db)
    psi = s.psi.copy()
    perf = sol.solve(psi, s.source)
    assert perf.solver_name == "MultigridcudaGKOCG"
    assert abs(perf.n_iterations - o.n_iterations) <= ITER_TOL and rel_l2(psi, o.x[0]) <= L2_TOL
    assert len(sol.ctx.mg_levels()) == 4
    sol = lduMatrix_solver_New("p", s, controls, db)
    psi2 = s.psi.copy()
    perf2 = sol.solve(psi2, s.source)
    assert sol.ctx.get_option("precond_setups") == 1 and perf2.n_iterations == perf.n_iterations
    assert np.array_equal(psi2, psi)
    with pytest.raises(FatalError):
        lduMatrix_solver_New("q", s, dict(controls, preconditioner=dict(pre, cycle="w")), ObjectRegistry())
    sol.ctx.close()
