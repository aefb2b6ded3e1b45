"""Short BiCGStab / GMRES / CG+BJ(4) runs for ncu (graphs off): python tools/ncu_target2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases, host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def setup(s):
    ctx = Context()
    ir, ic = host.collect_local_interface_indices(s)
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, s.symmetric, ir, ic)
    ctx.values_update(s.diag, s.upper, None if s.symmetric else s.lower, host.collect_interface_coeffs(s, True), None)
    ctx.vector_upload(L.OGL_VEC_B, s.source)
    ctx.vector_fill(L.OGL_VEC_X, 0.0)
    ctx.set_option("use_graph", 0)
    return ctx


which = sys.argv[1]
if which == "bicgstab":
    ctx = setup(cases.momentum_3d(200)[0])
    ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
    ctx.solve(L.OGL_SOLVER_BICGSTAB, tolerance=0.0, max_iter=12)
elif which == "bj4":
    ctx = setup(cases.pressure_3d(200)[0])
    ctx.precond_setup(L.OGL_PRECOND_BJ, 4)
    ctx.solve(L.OGL_SOLVER_CG, tolerance=0.0, max_iter=8)
else:
    ctx = setup(cases.channel((128, 64, 64), (1, 1, 1))[0])
    ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
    ctx.solve(L.OGL_SOLVER_GMRES, tolerance=0.0, max_iter=40, krylov_dim=100)
ctx.close()
