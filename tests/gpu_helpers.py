"""Shared helpers of the -m gpu parity tests: drive libogl_b200.so through the
C ABI exactly as the host layer does, next to the oracle on the same inputs."""
import numpy as np

from ogl_b200 import _lib as L
from ogl_b200 import host
from ogl_b200.backend import Context

SOLVER_ID = {"GKOCG": L.OGL_SOLVER_CG, "GKOBiCGStab": L.OGL_SOLVER_BICGSTAB,
             "GKOGMRES": L.OGL_SOLVER_GMRES}


PRECOND_ID = {"none": L.OGL_PRECOND_NONE, "BJ": L.OGL_PRECOND_BJ, "ISAI": L.OGL_PRECOND_ISAI,
              "GISAI": L.OGL_PRECOND_GISAI, "ILU": L.OGL_PRECOND_ILU, "IC": L.OGL_PRECOND_IC,
              "IRILU": L.OGL_PRECOND_IRILU, "Multigrid": L.OGL_PRECOND_MULTIGRID}


def upload_system(ctx: Context, s, scaling=1.0, partition=True):
    ir, ic = host.collect_local_interface_indices(s)
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, s.symmetric, ir, ic)
    if partition:
        ctx.partition_create(s.n, *host.create_communication_pattern(s))
        ctx.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
    ctx.values_update(s.diag, s.upper, None if s.symmetric else s.lower,
                      host.collect_interface_coeffs(s, True),
                      host.collect_interface_coeffs(s, False), scaling)
    ctx.vector_upload(L.OGL_VEC_B, s.source, scaling)
    ctx.vector_upload(L.OGL_VEC_X, s.psi)


def gpu_solve(ctx: Context, solver, precond, mbs=1, **kw):
    ctx.precond_setup(PRECOND_ID[precond], mbs)
    max_iter = kw.pop("max_iter", 1000)
    if solver == "GKOBiCGStab":
        max_iter *= 2   # StoppingCriterion.H:188, done by the host layer
    r = ctx.solve(SOLVER_ID[solver], max_iter=max_iter, **kw)
    return r, ctx.vector_download(L.OGL_VEC_X)


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
