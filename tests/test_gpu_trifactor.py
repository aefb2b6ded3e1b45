"""-m gpu: the `preconditioner` keywords ILU, IC and IRILU (Preconditioner.H:106-124,
143-176, 177-196) through the C ABI against the oracle's restatement (oracle/trifactor.hpp).

Bars: the incomplete factors and one application of the preconditioner are BIT-EXACT (the device
factorises level by level and sweeps in dependency order, but every row performs the sequential
algorithm's operations in its order); solves keep the north-star bars (iterations +-2, relative L2
difference <= 1e-8)."""
import numpy as np
import pytest

from gpu_helpers import PRECOND_ID, gpu_solve, rel_l2, upload_system
from ogl_b200 import cases
from ogl_b200.backend import Context, OglError
from ogl_b200.host import ObjectRegistry
from ogl_b200.plugin import lduMatrix_solver_New

pytestmark = pytest.mark.gpu

ITER_TOL = 2
L2_TOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    c = Context()
    yield c
    c.close()


def _csr(oracle, s):
    a = oracle.assemble(s)
    rp = np.zeros(s.n + 1, np.int32)
    np.cumsum(np.bincount(a.rows, minlength=s.n), out=rp[1:])
    return a, rp


def _unstructured(n=20000, extra=30000):
    """A random (unstructured) lduAddressing with an SPD M-matrix on it: rows of every length,
    hundreds of dependency levels of uneven size."""
    from conftest import random_ldu_mesh
    rng = np.random.default_rng(n)
    lower, upper = random_ldu_mesh(rng, n, extra)
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


SYSTEMS = {
    "momentum": lambda: cases.momentum_3d(20)[0],
    "pressure_spd": lambda: cases.pressure_3d(24, sign=-1.0)[0],
    "channel": lambda: cases.channel((16, 8, 8), (1, 1, 1))[0],     # cyclic couplings in x and z
    "unstructured": _unstructured,
}


@pytest.mark.parametrize("kind,system", [("ILU", "momentum"), ("IC", "pressure_spd"), ("ILU", "pressure_spd"),
                                         ("ILU", "channel"), ("IC", "unstructured"), ("ILU", "unstructured")])
def test_factors_and_apply_bit_exact(ctx, oracle, kind, system):
    s = SYSTEMS[system]()
    upload_system(ctx, s, partition=False)
    a, rp = _csr(oracle, s)
    f_ref = oracle.trifactor(kind, s.n, rp, a.cols, a.vals)
    ctx.precond_setup(PRECOND_ID[kind], 1)
    f = ctx.precond_factors_download()
    assert np.array_equal(f, f_ref)
    levels = ctx.get_option("tri_levels_lower")
    assert 1 <= levels <= s.n and ctx.get_option("tri_levels_upper") >= 1
    r = np.random.default_rng(11).standard_normal(s.n)
    z_ref = oracle.trifactor_apply(kind, s.n, rp, a.cols, f_ref, r)
    for variant in (0, 1):
        ctx.set_option("tri_variant", variant)
        z = ctx.precond_apply(r)
        assert np.array_equal(z, z_ref), f"tri_variant {variant}"
    ctx.set_option("tri_variant", 1)


def test_hex_mesh_levels_are_the_diagonal_planes(ctx):
    """7-point stencil on an N^3 box: row (i,j,k) depends on (i-1,j,k), (i,j-1,k), (i,j,k-1), so the
    dependency levels are the planes i + j + k = const: 3N - 2 of them, both directions."""
    n = 12
    s = cases.pressure_3d(n, sign=-1.0)[0]
    upload_system(ctx, s, partition=False)
    ctx.precond_setup(PRECOND_ID["IC"], 1)
    assert ctx.get_option("tri_levels_lower") == 3 * n - 2
    assert ctx.get_option("tri_levels_upper") == 3 * n - 2


def test_irilu_apply_bit_exact(ctx, oracle):
    s = cases.momentum_3d(16)[0]
    upload_system(ctx, s, partition=False)
    a, rp = _csr(oracle, s)
    f_ref = oracle.trifactor("IRILU", s.n, rp, a.cols, a.vals)
    ctx.precond_setup(PRECOND_ID["IRILU"], 1)
    assert np.array_equal(ctx.precond_factors_download(), f_ref)
    r = np.random.default_rng(12).standard_normal(s.n)
    assert np.array_equal(ctx.precond_apply(r), oracle.trifactor_apply("IRILU", s.n, rp, a.cols, f_ref, r))


def _check(ctx, oracle, s, solver, precond, **kw):
    upload_system(ctx, s, partition=False)
    r, x = gpu_solve(ctx, solver, precond, 1, **dict(kw))
    o = oracle.solve([oracle.assemble(s)], solver, precond, **kw)
    assert abs(r.n_iterations - o.n_iterations) <= ITER_TOL, (r.n_iterations, o.n_iterations)
    assert rel_l2(x, o.x[0]) <= L2_TOL
    assert r.init_residual == pytest.approx(o.init_residual, rel=1e-10)
    assert r.final_residual < kw.get("tolerance", 1e-6)
    return r, o


@pytest.mark.parametrize("solver,precond,system,kw", [
    ("GKOCG", "IC", "pressure_spd", {}),
    ("GKOCG", "ILU", "pressure_spd", {}),
    ("GKOBiCGStab", "ILU", "momentum", {}),
    ("GKOGMRES", "ILU", "momentum", {"krylov_dim": 30}),
    ("GKOBiCGStab", "IRILU", "momentum", {}),
    ("GKOGMRES", "IRILU", "momentum", {"krylov_dim": 30}),
    ("GKOGMRES", "ILU", "channel", {"krylov_dim": 30}),
    ("GKOCG", "IC", "unstructured", {}),
])
def test_solves_match_oracle(ctx, oracle, solver, precond, system, kw):
    s = SYSTEMS[system]()
    r, o = _check(ctx, oracle, s, solver, precond, tolerance=1e-9, **kw)
    upload_system(ctx, s, partition=False)
    rj, _ = gpu_solve(ctx, solver, "BJ", tolerance=1e-9, **kw)
    assert r.n_iterations < rj.n_iterations          # a stronger preconditioner than scalar Jacobi


def test_both_sweep_variants_give_the_same_solve(ctx):
    s = cases.pressure_3d(32, sign=-1.0)[0]
    out = []
    for variant in (0, 1):
        ctx.set_option("tri_variant", variant)
        upload_system(ctx, s, partition=False)
        r, x = gpu_solve(ctx, "GKOCG", "IC", tolerance=1e-8)
        out.append((r.n_iterations, r.final_residual, x))
    ctx.set_option("tri_variant", 1)
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2])


def test_graph_and_stream_paths_agree(ctx):
    s = cases.momentum_3d(16)[0]
    out = []
    for use_graph in (1, 0):
        ctx.set_option("use_graph", use_graph)
        upload_system(ctx, s, partition=False)
        r, x = gpu_solve(ctx, "GKOBiCGStab", "ILU", tolerance=1e-9)
        out.append((r.n_iterations, x))
    ctx.set_option("use_graph", 1)
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


def test_plugin_keywords_scaling_and_caching(oracle):
    """`preconditioner IC` needs an SPD matrix: OpenFOAM's pressure matrix is negative definite and
    takes `scaling -1` (README.md:101).  With `caching 2` the factors of the first solve serve the
    next two (Preconditioner.H:384-422)."""
    s = cases.pressure_3d(16)[0]
    spd = cases.pressure_3d(16, sign=-1.0)[0]
    o = oracle.solve([oracle.assemble(spd)], "GKOCG", "IC", tolerance=1e-9)
    controls = {"solver": "GKOCG", "executor": "cuda", "tolerance": 1e-9, "relTol": 0.0, "adaptMinIter": False,
                "scaling": -1.0, "preconditioner": {"preconditioner": "IC", "caching": 2}}
    db = ObjectRegistry()
    sol = lduMatrix_solver_New("p", s, controls, db)
    psi = s.psi.copy()
    perf = sol.solve(psi, s.source)
    assert perf.solver_name == "ICcudaGKOCG"
    assert abs(perf.n_iterations - o.n_iterations) <= ITER_TOL and rel_l2(psi, o.x[0]) <= L2_TOL
    for _ in range(2):
        sol = lduMatrix_solver_New("p", s, controls, db)
        psi = s.psi.copy()
        sol.solve(psi, s.source)
    assert sol.ctx.get_option("precond_setups") == 1
    sol = lduMatrix_solver_New("p", s, controls, db)
    sol.solve(s.psi.copy(), s.source)
    assert sol.ctx.get_option("precond_setups") == 2
    sol.ctx.close()


def test_regenerate_with_cached_factors():
    """`regenerate true` rebuilds the sparsity pattern every solve while `caching` keeps the factors:
    the dependency levels are re-analysed for the new pattern, the cached factors still apply."""
    s = cases.momentum_3d(12)[0]
    controls = {"solver": "GKOBiCGStab", "executor": "cuda", "tolerance": 1e-9, "relTol": 0.0, "adaptMinIter": False,
                "regenerate": True, "updateInitGuess": True, "preconditioner": {"preconditioner": "ILU", "caching": 3}}
    db = ObjectRegistry()
    its = []
    for _ in range(3):
        sol = lduMatrix_solver_New("U", s, controls, db)
        psi = s.psi.copy()
        its.append(sol.solve(psi, s.source).n_iterations)
        true_res = np.abs(sol.ctx.spmv(psi) - s.source).sum() / sol.last_result.norm_factor
        assert true_res < 1e-9
    assert its[0] == its[1] == its[2] and sol.ctx.get_option("precond_setups") == 1
    sol.ctx.close()


def test_rejections(ctx):
    s = cases.momentum_3d(8)[0]
    upload_system(ctx, s, partition=False)
    with pytest.raises(OglError):
        ctx.precond_setup(PRECOND_ID["IC"], 1)          # IC on an asymmetric matrix
    with pytest.raises(OglError):
        ctx.precond_factors_download()                  # nothing was set up
    ctx.precond_setup(PRECOND_ID["BJ"], 1)
    with pytest.raises(OglError):
        ctx.precond_factors_download()
