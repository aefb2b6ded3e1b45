"""Timings of the BASELINE.json configs that are parity cases rather than the bench line:
    configs[2] 200^3 momentum GKOBiCGStab+BJ, GMRES on a channel box, BJ block sizes.
Prints one JSON line per case (device-resident solve, CUDA-event time of the loop)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases, host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def run(name, s, solver, precond, mbs, tol, reps=3, **kw):
    ctx = Context()
    ir, ic = host.collect_local_interface_indices(s)
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, s.symmetric, ir, ic)
    ctx.values_update(s.diag, s.upper, None if s.symmetric else s.lower,
                      host.collect_interface_coeffs(s, True), None)
    ctx.vector_upload(L.OGL_VEC_B, s.source)
    best = None
    for _ in range(reps):
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.precond_setup(L.OGL_PRECOND_BJ if precond == "BJ" else L.OGL_PRECOND_NONE, mbs)
        r = ctx.solve(solver, tolerance=tol, rel_tol=0.0, **kw)
        if best is None or r.solve_us < best.solve_us:
            best = r
    n, nnz = s.n, ctx.nnz
    per_iter = {L.OGL_SOLVER_CG: 12 * nnz + 4 * (n + 1) + 96 * n,
                L.OGL_SOLVER_BICGSTAB: 2 * (12 * nnz + 4 * (n + 1)) + 200 * n}.get(solver)
    it = max(best.n_iterations, 1)
    row = dict(case=name, rows=n, nnz=nnz, iterations=best.n_iterations,
               final_residual=best.final_residual, solve_ms=round(best.solve_us / 1e3, 3),
               us_per_iteration=round(best.solve_us / it, 2), it_per_s=round(it / best.solve_us * 1e6, 1))
    if per_iter:
        row["alg_GBs"] = round(per_iter * it / best.solve_us / 1e3, 1)
    x = ctx.vector_download(L.OGL_VEC_X)
    row["err_vs_manufactured"] = float(np.linalg.norm(x - s.x_star) / np.linalg.norm(s.x_star))
    print(json.dumps(row), flush=True)
    ctx.close()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    t = time.time()
    mom = cases.momentum_3d(n)[0]
    run(f"configs[2] {n}^3 momentum GKOBiCGStab+BJ", mom, L.OGL_SOLVER_BICGSTAB, "BJ", 1, 1e-5, max_iter=2000)
    run(f"{n}^3 momentum GKOBiCGStab+BJ(4)", mom, L.OGL_SOLVER_BICGSTAB, "BJ", 4, 1e-5, max_iter=2000)
    run(f"{n}^3 momentum GKOGMRES(30)+BJ", mom, L.OGL_SOLVER_GMRES, "BJ", 1, 1e-5, krylov_dim=30)
    pre = cases.pressure_3d(n)[0]
    run(f"{n}^3 pressure GKOCG+BJ(4)", pre, L.OGL_SOLVER_CG, "BJ", 4, 1e-6)
    run(f"{n}^3 pressure GKOCG none", pre, L.OGL_SOLVER_CG, "none", 1, 1e-6)
    ch = cases.channel((2 * n // 2, n // 2, n // 2), (1, 1, 1))[0]
    run(f"channel {2*n//2}x{n//2}x{n//2} cyclic GKOGMRES(100)+BJ", ch, L.OGL_SOLVER_GMRES, "BJ", 1, 1e-6,
        krylov_dim=100)
    print(json.dumps({"wall_s": round(time.time() - t, 1)}))


if __name__ == "__main__":
    main()
