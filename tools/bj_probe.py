import sys, os
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bench.py") else ".")
import bench
from ogl_b200 import _lib as L
from ogl_b200.backend import Context
s = bench.build_rank_system(200, 1, 0)
ctx = Context()
ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
ctx.values_update(s.diag, s.upper)
ctx.vector_upload(L.OGL_VEC_B, s.source)
for mbs in (1, 2, 4, 8):
    best = None
    for _ in range(3):
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.precond_setup(L.OGL_PRECOND_BJ, mbs)
        r = ctx.solve(L.OGL_SOLVER_CG, tolerance=1e-6, max_iter=1000)
        if best is None or r.solve_us < best.solve_us: best = r
    print({"mbs": mbs, "iters": best.n_iterations, "us_per_iter": round(best.solve_us / best.n_iterations, 1), "solve_ms": round(best.solve_us/1e3, 2)}, flush=True)
