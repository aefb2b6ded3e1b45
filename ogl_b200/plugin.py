"""The plugin surface: GKOCG / GKOBiCGStab / GKOGMRES behind an
`lduMatrix::solver::New`-style selector, with the fvSolution keyword semantics
of the reference.

    Solver/CG/GKOCG.{H,C}, Solver/BiCGStab/GKOBiCGStab.{H,C}, Solver/GMRES/GKOGMRES.{H,C}
    BaseWrapper/lduBase/GKOlduBase.H:23-62, lduLduBase/lduLduBase.H:189-332
    StoppingCriterion/StoppingCriterion.H:164-234, Preconditioner/Preconditioner.H:353-431
    common/common.C:75-146 (per-field key-value store)

This Python layer exists for the tests and the benchmark; the C++ host layer in
ogl_b200/host_cpp mirrors the same classes for an OpenFOAM build.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Type

import numpy as np

from . import _lib as L
from .backend import Context
from .cases import LduSystem
from .host import FatalError, HostMatrixWrapper, ObjectRegistry
from .parallel import Pstream


# `preconditioner` keyword (Preconditioner.H:91-105 BJ, :106-124 ILU, :143-176 IRILU, :177-196 IC,
# :225-242 ISAI, :243-260 GISAI, :261-341 Multigrid); ILUT / ICT (ParILUT / ParICT) are rejected
PRECOND_KINDS = {"none": L.OGL_PRECOND_NONE, "BJ": L.OGL_PRECOND_BJ, "ISAI": L.OGL_PRECOND_ISAI,
                 "GISAI": L.OGL_PRECOND_GISAI, "ILU": L.OGL_PRECOND_ILU, "IC": L.OGL_PRECOND_IC,
                 "IRILU": L.OGL_PRECOND_IRILU, "Multigrid": L.OGL_PRECOND_MULTIGRID}


@dataclass
class SolverPerformance:
    solver_name: str
    field_name: str
    initial_residual: float = 0.0
    final_residual: float = 0.0
    n_iterations: int = 0


# ---- per-field key-value store (common/common.C:75-146) --------------------

def _props(db: ObjectRegistry, field: str) -> dict:
    return db.setdefault(field + "_gkoSolverProperties", {})


def set_gko_solver_property(field, db, key, value):
    # common.C:75-76: the value travels through a `label` parameter -> truncation
    _props(db, field)[key] = int(value)


def get_gko_solver_property(field, key, db, default):
    # common.C:92-102: read back into a `label`
    if db.found_object(field + "_gkoSolverProperties"):
        return int(_props(db, field).get(key, default))
    return default


def get_solve_prev_iters(field, db, is_final):
    return get_gko_solver_property(field, "prevSolveIters_final" if is_final else "prevSolveIters",
                                   db, 1)


def set_solve_prev_iters(field, db, iters, is_final):
    set_gko_solver_property(field, db, "prevSolveIters_final" if is_final else "prevSolveIters", iters)


def get_solve_prev_rel_res_cost(field, db):
    return get_gko_solver_property(field, "_prev_solve", db, 0.0)


def set_solve_prev_rel_res_cost(field, db, cost):
    set_gko_solver_property(field, db, "_prev_solve", cost if math.isfinite(cost) else 0)


# ---- stopping criterion ------------------------------------------------------

class StoppingCriterion:
    """StoppingCriterion.H:164-234 (keyword defaults and the adaptive
    minIter / evaluation frequency)."""

    def __init__(self, controls: dict):
        self.max_iter = int(controls.get("maxIter", 1000))
        self.min_iter = int(controls.get("minIter", 0))
        self.tolerance = float(controls.get("tolerance", 1e-6))
        self.rel_tol = float(controls.get("relTol", 1e-6))
        self.norm_eval_limit = int(controls.get("normEvalLimit", 100))
        self.frequency = int(controls.get("evalFrequency", 1))
        self.relaxation = float(controls.get("relaxationFactor", 0.6))
        self.adapt_min_iter = bool(controls.get("adaptMinIter", True))
        if controls.get("solver") == "GKOBiCGStab":
            self.max_iter *= 2   # :188

    @property
    def is_final(self) -> bool:
        return self.rel_tol == 0.0

    def effective(self, export_res: bool, prev_solve_iters: int, prev_rel_cost: float):
        min_iter, frequency = self.min_iter, self.frequency
        if not export_res and prev_solve_iters > 0 and self.adapt_min_iter and prev_rel_cost > 0:
            min_iter = int(prev_solve_iters * self.relaxation)
            alpha = math.sqrt(1.0 / (prev_solve_iters * (1.0 - self.relaxation)) * prev_rel_cost)
            frequency = min(self.norm_eval_limit, max(1, int(1 / alpha)))
        return min_iter, frequency


# ---- solver base ---------------------------------------------------------------

class GKOlduBaseSolver:
    type_name = "GKObase"
    solver_id = -1
    symmetric_only = False

    def __init__(self, field_name: str, matrix: LduSystem, controls: dict,
                 db: ObjectRegistry, pstream: Optional[Pstream] = None):
        self.field_name, self.matrix, self.controls, self.db = field_name, matrix, controls, db
        self.pstream = pstream or Pstream()
        self.verbose = int(controls.get("verbose", 0))
        self.criterion = StoppingCriterion(controls)
        # ExecutorHandler (ExecutorHandler.H:125-147): registry name "<executor>_<field>"
        self.exec_name = str(controls.get("executor", "reference"))
        if self.exec_name != "cuda":
            raise FatalError(f"OGL-B200 does not support the executor: {self.exec_name}\n"
                             "Valid choices are: cuda (this backend has no CPU executor)")
        ranks_per_gpu = int(controls.get("ranksPerGPU", 1))
        if ranks_per_gpu != 1:
            raise FatalError("ranksPerGPU != 1 is not implemented (as in the reference, "
                             "Partition.H:69-70)")
        key = f"{self.exec_name}_{field_name}"
        if not db.found_object(key):
            db[key] = Context(device_id=self.pstream.local_rank // ranks_per_gpu,
                              rank=self.pstream.rank, n_ranks=self.pstream.n_ranks,
                              nccl_id=self.pstream.nccl_id)
        self.ctx: Context = db[key]
        # matrixFormat (lduLduBase.H:55-56, CsrMatrixWrapper.H:249-250): Coo (the reference's
        # default) and Csr both map to the CSR kernels -- same row-major order --, Ell to the ELL copy
        fmt = str(controls.get("matrixFormat", "Coo"))
        if fmt not in ("Coo", "Csr", "Ell"):
            raise FatalError(f"unknown matrixFormat {fmt}\nValid choices are: Coo, Csr, Ell")
        # one set_option per key with its final value: an unchanged value keeps the cached chunk graph
        opts = {"spmv_variant": 7 if fmt == "Ell" else 0}
        for opt in ("spmv_variant", "chunk_iters", "use_graph", "comm_mode", "fused_halo", "ghost_p",
                    "fused_pcg", "device_loop", "loop_iters", "l2_keep_mb", "ell_auto", "fuse_p", "ell_coded", "ell_tma", "gmres_persist", "tri_variant"):
            if opt in controls:
                opts[opt] = int(controls[opt])
        for opt, val in opts.items():
            self.ctx.set_option(opt, val)
        # preconditioner keyword: word or sub-dict (Preconditioner.H:363-382)
        pre = controls.get("preconditioner")
        if pre is None:
            raise FatalError("keyword preconditioner is undefined")
        self.precond_controls = pre if isinstance(pre, dict) else {}
        self.precond_name = pre["preconditioner"] if isinstance(pre, dict) else str(pre)
        if self.precond_name not in PRECOND_KINDS:
            raise FatalError(f"OGL does not support the preconditioner: {self.precond_name}\n"
                             "Valid Choices: none, BJ, ILU, IRILU, IC, ISAI, GISAI, Multigrid")
        if self.precond_name in ("ISAI", "GISAI") and int(self.precond_controls.get("sparsityPower", 1)) != 1:
            raise FatalError("ISAI / GISAI: only sparsityPower 1 is implemented")
        if self.precond_name == "Multigrid":
            # Preconditioner.H:297-320: cycle (v | w | f), maxLevels, minCoarseRows, coarseSolverIters
            if str(self.precond_controls.get("cycle", "v")) != "v":
                raise FatalError("Multigrid: only cycle v is implemented")
            self.ctx.set_option("mg_max_levels", int(self.precond_controls.get("maxLevels", 9)))
            self.ctx.set_option("mg_min_coarse_rows", int(self.precond_controls.get("minCoarseRows", 10)))
            self.ctx.set_option("mg_coarse_iters", int(self.precond_controls.get("coarseSolverIters", 4)))
        self.host_matrix = HostMatrixWrapper(db, matrix, controls, field_name, self.ctx, self.pstream)

    # lduLduBase.H:189-308
    def solve(self, psi: np.ndarray, source: np.ndarray) -> SolverPerformance:
        c, ctx, db, f = self.controls, self.ctx, self.db, self.field_name
        perf = SolverPerformance(self.precond_name + self.exec_name + self.type_name, f)
        scaling = float(c.get("scaling", 1.0))
        # PersistentVector b: re-uploaded when updateRHS (default true), :217-226
        if not db.found_object(f + "_rhs") or c.get("updateRHS", True):
            ctx.vector_upload(L.OGL_VEC_B, source, scaling)   # :242-252 scale_RHS
            db[f + "_rhs"] = True
        # PersistentVector x: kept from the previous solve unless updateInitGuess, :228-237
        if not db.found_object(f + "_solution") or c.get("updateInitGuess", False):
            ctx.vector_upload(L.OGL_VEC_X, psi)
            db[f + "_solution"] = True
        # preconditioner: regenerated every solve unless cached (Preconditioner.H:384-422)
        cache = get_gko_solver_property(f, "preconditionerCaching", db, 0)
        if db.found_object("Cached_preconditioner_" + f) and cache > 0:
            set_gko_solver_property(f, db, "preconditionerCaching", cache - 1)
        else:
            set_gko_solver_property(f, db, "preconditionerCaching",
                                    int(self.precond_controls.get("caching", 0)))
            ctx.precond_setup(PRECOND_KINDS[self.precond_name],
                              int(self.precond_controls.get("maxBlockSize", 1)),
                              bool(self.precond_controls.get("skipSorting", True)))
            db["Cached_preconditioner_" + f] = True
        if c.get("debug", False) and c.get("writeTime", False):
            self.write()   # lduLduBase.H:259-264
        export_res = bool(c.get("export", False))
        crit = self.criterion
        min_iter, frequency = crit.effective(
            export_res, get_solve_prev_iters(f, db, crit.is_final),
            get_solve_prev_rel_res_cost(f, db))
        self.last_min_iter, self.last_frequency = min_iter, frequency
        res = ctx.solve(self.solver_id, crit.tolerance, crit.rel_tol, min_iter, crit.max_iter,
                        frequency, int(c.get("krylovDim", 100)), export_res)
        ctx.vector_download(L.OGL_VEC_X, psi)   # copy_back, Vector.H:144-167
        perf.initial_residual = res.init_residual
        perf.final_residual = res.final_residual
        perf.n_iterations = res.n_iterations
        # lduLduBase.H:286-293
        set_solve_prev_iters(f, db, res.criterion_calls, crit.is_final)
        t_iter = res.solve_us / max(res.n_iterations, 1)
        rel_cost = t_iter / res.resnorm_us if res.resnorm_us > 0 else 0.0
        if self.pstream.par_run:
            # every rank must derive the same minIter / frequency next time: rank 0's figure
            # (lduLduBase.H:290-291 broadcasts it over the host communicator)
            rel_cost = self.pstream.broadcast_scalar(rel_cost)
        set_solve_prev_rel_res_cost(f, db, rel_cost)
        self.last_result = res
        return perf

    # CsrMatrixWrapper.H:273-290, Vector.H:173-176, common.C:31-58
    def write(self):
        import os
        folder = self.db.time_path()
        os.makedirs(folder, exist_ok=True)
        f = self.field_name
        self.ctx.export_mtx(0, os.path.join(folder, f + "_A_local.mtx"))
        self.ctx.export_mtx(1, os.path.join(folder, f + "_A_non_local.mtx"))
        self.ctx.export_mtx(2, os.path.join(folder, f + "_rhs_b_.mtx"))
        # side-car with the communication pattern, so that a decomposed dump can be read back
        # (mtxio.import_decomposed); the reference's files do not carry it
        self.ctx.export_mtx(3, os.path.join(folder, f + "_partition.json"))


class GKOCG(GKOlduBaseSolver):
    type_name = "GKOCG"
    solver_id = L.OGL_SOLVER_CG
    symmetric_only = True     # GKOCG.C:16-17: addsymMatrixConstructorToTable only


class GKOBiCGStab(GKOlduBaseSolver):
    type_name = "GKOBiCGStab"
    solver_id = L.OGL_SOLVER_BICGSTAB   # GKOBiCGStab.C:16-20: sym + asym


class GKOGMRES(GKOlduBaseSolver):
    type_name = "GKOGMRES"
    solver_id = L.OGL_SOLVER_GMRES      # GKOGMRES.C:16-20: sym + asym


_SYM_TABLE: Dict[str, Type[GKOlduBaseSolver]] = {c.type_name: c for c in (GKOCG, GKOBiCGStab, GKOGMRES)}
_ASYM_TABLE: Dict[str, Type[GKOlduBaseSolver]] = {c.type_name: c for c in (GKOBiCGStab, GKOGMRES)}


def lduMatrix_solver_New(field_name: str, matrix: LduSystem, controls: dict, db: ObjectRegistry,
                         pstream: Optional[Pstream] = None) -> GKOlduBaseSolver:
    """lduMatrix::solver::New: run-time selection by the `solver` keyword in the
    symmetric or asymmetric constructor table."""
    name = controls.get("solver")
    table = _SYM_TABLE if matrix.symmetric else _ASYM_TABLE
    if name not in table:
        kind = "symmetric" if matrix.symmetric else "asymmetric"
        raise FatalError(f"Unknown {kind} matrix solver {name}\nValid {kind} matrix solvers are : "
                         + " ".join(sorted(table)))
    return table[name](field_name, matrix, controls, db, pstream)
