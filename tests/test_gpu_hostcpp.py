"""-m gpu: the C++ plugin layer (lduMatrix::solver::New -> GKO* -> lduLduBase ->
HostMatrixWrapper -> C ABI) against the oracle, read like the reference would
be driven by OpenFOAM."""
import numpy as np
import pytest

from foam_harness import FoamCase, FoamFatalError
from gpu_helpers import rel_l2
from ogl_b200 import cases

pytestmark = pytest.mark.gpu

BASE = {"executor": "cuda", "relTol": 0.0, "adaptMinIter": False}


@pytest.mark.parametrize("builder,solver,precond,tol", [
    (lambda: cases.pressure_3d(20)[0], "GKOCG", "BJ", 1e-9),
    (lambda: cases.pressure_3d(16)[0], "GKOCG", "none", 1e-9),
    (lambda: cases.pressure_3d(16)[0], "GKOCG", {"preconditioner": "BJ", "maxBlockSize": 4}, 1e-9),
    (lambda: cases.momentum_3d(16)[0], "GKOBiCGStab", "BJ", 1e-10),
    (lambda: cases.channel((16, 8, 8), (1, 1, 1))[0], "GKOGMRES", "BJ", 1e-8),
    (lambda: cases.pressure_3d(16, sign=-1.0)[0], "GKOCG", {"preconditioner": "ISAI", "sparsityPower": 1}, 1e-9),
    (lambda: cases.momentum_3d(14)[0], "GKOBiCGStab", "GISAI", 1e-10),
    (lambda: cases.pressure_3d(16, sign=-1.0)[0], "GKOCG", "IC", 1e-9),
    (lambda: cases.momentum_3d(14)[0], "GKOBiCGStab", "ILU", 1e-10),
    (lambda: cases.momentum_3d(14)[0], "GKOGMRES", "IRILU", 1e-10),
    (lambda: cases.momentum_3d(14)[0], "GKOBiCGStab", "Multigrid", 1e-10),
])
def test_plugin_solve_matches_oracle(oracle, builder, solver, precond, tol):
    s = builder()
    c = FoamCase(s)
    try:
        controls = dict(BASE, solver=solver, preconditioner=precond, tolerance=tol, krylovDim=30)
        psi, name, r0, r1, it = c.solve("f", controls, s.psi, s.source)
        pname = precond if isinstance(precond, str) else precond["preconditioner"]
        mbs = 1 if isinstance(precond, str) else precond.get("maxBlockSize", 1)
        assert name == f"{pname}cuda{solver}"
        o = oracle.solve([oracle.assemble(s)], solver, pname, max_block_size=mbs, tolerance=tol,
                         krylov_dim=30)
        assert abs(it - o.n_iterations) <= 2
        assert rel_l2(psi, o.x[0]) <= 1e-8
        assert r0 == pytest.approx(o.init_residual, rel=1e-10) and r1 < tol
    finally:
        c.close()


def test_time_steps_reuse_the_registry(oracle):
    s = cases.pressure_3d(16)[0]
    c = FoamCase(s)
    try:
        controls = dict(BASE, solver="GKOCG", preconditioner="BJ", tolerance=1e-9)
        psi1, *_ , it1 = c.solve("p", controls, s.psi, s.source)
        n_objects = c.registry_size()
        # second time step: new coefficients (scaled matrix), new rhs; initial guess is the
        # previous DEVICE solution whatever psi says (updateInitGuess false)
        s2 = cases.pressure_3d(16)[0]
        s2.diag, s2.upper = 2.0 * s.diag, 2.0 * s.upper
        c.set_coeffs(s2)
        psi2, _, _, _, it2 = c.solve("p", controls, np.full(s.n, 7.0), 3.0 * s.source)
        assert c.registry_size() == n_objects          # nothing re-created
        a = oracle.assemble(s2)
        a.b = 3.0 * a.b
        a.x = psi1.copy()
        o = oracle.solve([a], "GKOCG", "BJ", tolerance=1e-9)
        assert abs(it2 - o.n_iterations) <= 2 and rel_l2(psi2, o.x[0]) <= 1e-8
        # updateInitGuess true: psi is uploaded again
        psi3, _, _, _, it3 = c.solve("p", dict(controls, updateInitGuess=True), np.zeros(s.n),
                                     3.0 * s.source)
        a.x = np.zeros(s.n)
        o3 = oracle.solve([a], "GKOCG", "BJ", tolerance=1e-9)
        assert abs(it3 - o3.n_iterations) <= 2 and rel_l2(psi3, o3.x[0]) <= 1e-8
    finally:
        c.close()


def test_unsupported_interfaces_and_preconditioners():
    s = cases.channel((8, 4, 4), (1, 1, 1))[0]
    s.interfaces[-1].kind = "cyclicAMI"
    c = FoamCase(s)
    try:
        with pytest.raises(FoamFatalError, match="CyclicAMIFvPatch"):
            c.solve("p", dict(BASE, solver="GKOCG", preconditioner="BJ"), s.psi, s.source)
    finally:
        c.close()
    s = cases.pressure_3d(6)[0]
    c = FoamCase(s)
    try:
        with pytest.raises(FoamFatalError, match="does not support the preconditioner: ILUT"):
            c.solve("p", dict(BASE, solver="GKOCG", preconditioner="ILUT"), s.psi, s.source)
    finally:
        c.close()


def test_caching_regenerate_and_adaptive_defaults_through_the_cpp_layer(oracle):
    """The keyword paths the advisor flagged in round 1, driven through the C++ classes: `caching N`
    reuses the preconditioner, `regenerate true` keeps the device vectors, and with the reference's
    default keywords the adaptive minIter is live from the second solve on."""
    s = cases.pressure_3d(14)[0]
    c = FoamCase(s)
    try:
        controls = dict(BASE, solver="GKOCG", tolerance=1e-9, updateInitGuess=True,
                        preconditioner={"preconditioner": "BJ", "caching": 2})
        its = [c.solve("p", controls, np.zeros(s.n), s.source)[4] for _ in range(4)]
        assert len(set(its)) == 1 and its[0] > 10          # same matrix: cached or not, same solve
        psi, *_, it1 = c.solve("q", dict(BASE, solver="GKOCG", preconditioner="BJ", tolerance=1e-9,
                                         regenerate=True), np.zeros(s.n), s.source)
        psi2, *_, it2 = c.solve("q", dict(BASE, solver="GKOCG", preconditioner="BJ", tolerance=1e-9,
                                          regenerate=True), np.zeros(s.n), s.source)
        assert it1 > 10 and it2 <= 2 and rel_l2(psi2, psi) < 1e-7
        # defaults (adaptMinIter true): the second solve may not stop before 0.6 x the first one's calls
        dflt = {"executor": "cuda", "solver": "GKOCG", "preconditioner": "BJ", "tolerance": 1e-9, "relTol": 0.0}
        _, _, _, _, a1 = c.solve("r", dflt, np.zeros(s.n), s.source)
        # new right-hand side, x0 = previous device solution: whatever it would need, not fewer than minIter
        _, _, _, _, a2 = c.solve("r", dflt, np.zeros(s.n), 1.5 * s.source)
        assert a1 > 10 and a2 >= int(0.6 * a1)
    finally:
        c.close()
