"""TEST INFRASTRUCTURE -- ctypes front end of liboracle.so / _ref/libogl_ref.so.

`assemble()` walks one rank's LduSystem through the same steps as the
reference's HostMatrixWrapper constructor (HostMatrix/HostMatrix.C:16-96):
count interface nnz (:158-178), communication pattern (:251-306), local
sparsity incl. cyclic merge (:468-589), non-local sparsity (:438-466), value
update (:592-732).  `solve()` runs the restated Ginkgo reference-executor
solvers under OGL's criterion on all ranks of a case at once.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/libogl_ref.so when /root/reference exists)."""
    lib = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("assembly.cpp", "krylov.cpp", "oracle.h")]
    stale = force or not os.path.exists(lib) or any(
        os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs)
    need_ref = (not os.path.exists(os.path.join(_HERE, "_ref", "libogl_ref.so"))
                and os.path.exists("/root/reference/HostMatrix/HostMatrixFreeFunctions.C"))
    if stale or need_ref or force:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       capture_output=True)


def _ip(a):
    return a.ctypes.data_as(i32p)


def _fp(a):
    return a.ctypes.data_as(f64p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RankSystem(C.Structure):
    _fields_ = [("n", C.c_int32), ("nnz", C.c_int32), ("rows", i32p), ("cols", i32p),
                ("vals", f64p), ("n_halo", C.c_int32), ("nl_rows", i32p), ("nl_cols", i32p),
                ("nl_vals", f64p), ("n_targets", C.c_int32), ("target_ids", i32p),
                ("target_sizes", i32p), ("send_idxs", i32p), ("b", f64p), ("x", f64p)]


class SolveParams(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("max_block_size", C.c_int32),
                ("tolerance", C.c_double), ("rel_tol", C.c_double), ("min_iter", C.c_int32),
                ("max_iter", C.c_int32), ("frequency", C.c_int32), ("krylov_dim", C.c_int32),
                ("threads", C.c_int32), ("mg_max_levels", C.c_int32), ("mg_min_coarse_rows", C.c_int32),
                ("mg_coarse_iters", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("init_residual", C.c_double), ("final_residual", C.c_double),
                ("criterion_calls", C.c_int32), ("n_iterations", C.c_int32),
                ("norm_factor", C.c_double), ("n_history", C.c_int32), ("seconds", C.c_double)]


SOLVERS = {"GKOCG": 0, "GKOBiCGStab": 1, "GKOGMRES": 2}
PRECONDS = {"none": 0, "BJ": 1, "ISAI": 2, "GISAI": 3, "ILU": 4, "IC": 5, "IRILU": 6, "Multigrid": 7}


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _LIB.orc_solve.restype = C.c_int
        _LIB.orc_bj_find_blocks.restype = C.c_int32
        _LIB.orc_time_spmv.restype = C.c_double
        _LIB.orc_time_spmv.argtypes = [C.c_int32, i32p, i32p, f64p, f64p, f64p, C.c_int, C.c_int]
    return _LIB


def ref_lib():
    """The reference's own free functions (oracle/_ref); None if not built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libogl_ref.so")
        if not os.path.exists(path):
            try:
                build()
            except Exception:
                pass
        if os.path.exists(path):
            _REF = C.CDLL(path)
    return _REF


# ---------------------------------------------------------------- assembly

def init_local_sparsity(nrows, upper_addr, lower_addr, symmetric, which="oracle"):
    upper_addr, lower_addr = _i32(upper_addr), _i32(lower_addr)
    F = upper_addr.size
    nnz = nrows + 2 * F
    rows, cols, perm = (np.zeros(nnz, np.int32) for _ in range(3))
    if which == "oracle":
        lib().orc_init_local_sparsity(C.c_int32(nrows), C.c_int32(F), C.c_int(int(symmetric)),
                                      _ip(upper_addr), _ip(lower_addr), _ip(rows), _ip(cols),
                                      _ip(perm))
    else:
        ref_lib().ref_init_local_sparsity(C.c_int(nrows), C.c_int(F), C.c_int(int(symmetric)),
                                          _ip(upper_addr), _ip(lower_addr), _ip(rows),
                                          _ip(cols), _ip(perm))
    return rows, cols, perm


def update_host(kind, permute, scale, diag, upper, lower=None, iface=None, which="oracle"):
    """The four host update free functions (HostMatrixFreeFunctions.C:21-102)."""
    permute, diag, upper = _i32(permute), _f64(diag), _f64(upper)
    total, F, nd = permute.size, upper.size, diag.size
    out = np.zeros(total)
    L = lib() if which in ("oracle", "as_written") else ref_lib()
    pre = "orc_" if which in ("oracle", "as_written") else "ref_"
    sc = C.c_double(scale)
    if kind == "symmetric":
        name = "orc_symmetric_update_as_written" if which == "as_written" else pre + "symmetric_update"
        getattr(L, name)(C.c_int32(total), C.c_int32(F), _ip(permute), sc, _fp(diag), _fp(upper),
                         _fp(out))
    elif kind == "non_symmetric":
        lower = _f64(lower)
        getattr(L, pre + "non_symmetric_update")(C.c_int32(total), C.c_int32(F), _ip(permute), sc,
                                                 _fp(diag), _fp(upper), _fp(lower), _fp(out))
    elif kind == "symmetric_w_interface":
        iface = _f64(iface)
        getattr(L, pre + "symmetric_update_w_interface")(
            C.c_int32(total), C.c_int32(nd), C.c_int32(F), _ip(permute), sc, _fp(diag), _fp(upper),
            _fp(iface), _fp(out))
    elif kind == "non_symmetric_w_interface":
        lower, iface = _f64(lower), _f64(iface)
        getattr(L, pre + "non_symmetric_update_w_interface")(
            C.c_int32(total), C.c_int32(nd), C.c_int32(F), _ip(permute), sc, _fp(diag), _fp(upper),
            _fp(lower), _fp(iface), _fp(out))
    else:
        raise ValueError(kind)
    return out


@dataclass
class Assembled:
    """What HostMatrixWrapper holds after construction, for one rank."""
    n: int
    rows: np.ndarray
    cols: np.ndarray
    ldu_mapping: np.ndarray
    vals: np.ndarray
    nl_rows: np.ndarray
    nl_cols: np.ndarray
    nl_mapping: np.ndarray
    nl_vals: np.ndarray
    target_ids: np.ndarray
    target_sizes: np.ndarray
    send_idxs: np.ndarray
    b: np.ndarray
    x: np.ndarray

    @property
    def row_ptrs(self):
        rp = np.zeros(self.n + 1, np.int32)
        np.cumsum(np.bincount(self.rows, minlength=self.n), out=rp[1:])
        return rp


def assemble(s, scaling: float = 1.0) -> Assembled:
    L = lib()
    n, F = s.n, s.n_faces
    sym = s.symmetric
    proc = [i for i in s.interfaces if i.kind == "processor"]
    cyc_idx = [k for k, i in enumerate(s.interfaces) if i.kind != "processor"]
    # collect_local_interface_indices (HostMatrix.C:385-410): rows = faceCells,
    # cols = patchAddr(neighbPatchID)
    if cyc_idx:
        ir = np.concatenate([s.interfaces[k].face_cells for k in cyc_idx]).astype(np.int32)
        ic = np.concatenate([s.interfaces[s.interfaces[k].nbr_patch].face_cells
                             for k in cyc_idx]).astype(np.int32)
        icoef = np.concatenate([s.interfaces[k].bou_coeffs for k in cyc_idx])
    else:
        ir = ic = np.zeros(0, np.int32)
        icoef = np.zeros(0)
    n_if = ir.size
    nnz = n + 2 * F + n_if
    rows, cols, perm = (np.zeros(nnz, np.int32) for _ in range(3))
    up, lo = _i32(s.upper_addr), _i32(s.lower_addr)
    L.orc_init_local_sparsity(C.c_int32(n), C.c_int32(F), C.c_int(int(sym)), _ip(up), _ip(lo),
                              _ip(rows), _ip(cols), _ip(perm))
    if n_if:
        L.orc_merge_local_interfaces(C.c_int32(n), C.c_int32(F), C.c_int(int(sym)),
                                     C.c_int32(n_if), _ip(ir), _ip(ic), _ip(rows), _ip(cols),
                                     _ip(perm))
    # staging [upper | lower | diag | -iface] then gather (HostMatrix.C:634-704)
    neg_if = np.zeros(n_if)
    if n_if:
        L.orc_negate(C.c_int32(n_if), _fp(_f64(icoef)), _fp(neg_if))
    parts = [_f64(s.upper)] + ([] if sym else [_f64(s.lower)]) + [_f64(s.diag), neg_if]
    staging = np.concatenate(parts)
    vals = np.zeros(nnz)
    L.orc_gather_from_staging(C.c_int32(nnz), _ip(perm), _fp(staging), _fp(vals))
    if scaling != 1.0:
        vals *= scaling   # documented intent (README.md:81), SURVEY Appendix B-3
    # communication pattern
    npi = len(proc)
    nbr = np.array([p.nbr_rank for p in proc], np.int32)
    sz = np.array([p.face_cells.size for p in proc], np.int32)
    fcs = (np.concatenate([p.face_cells for p in proc]).astype(np.int32) if npi
           else np.zeros(0, np.int32))
    n_halo = int(fcs.size)
    nt = C.c_int32(0)
    tid, tsz = np.zeros(max(npi, 1), np.int32), np.zeros(max(npi, 1), np.int32)
    sidx = np.zeros(max(n_halo, 1), np.int32)
    L.orc_comm_pattern(C.c_int32(npi), _ip(nbr), _ip(sz), _ip(fcs), C.byref(nt), _ip(tid),
                       _ip(tsz), _ip(sidx))
    # non-local pattern + values
    nlr, nlc, nlp = (np.zeros(max(n_halo, 1), np.int32) for _ in range(3))
    L.orc_non_local_pattern(C.c_int32(n_halo), _ip(fcs), _ip(nlr), _ip(nlc), _ip(nlp))
    bou = (np.concatenate([p.bou_coeffs for p in proc]) if npi else np.zeros(0))
    neg = np.zeros(max(n_halo, 1))
    L.orc_negate(C.c_int32(n_halo), _fp(_f64(bou) if n_halo else neg), _fp(neg))
    nlv = np.zeros(max(n_halo, 1))
    L.orc_non_local_update(C.c_int32(n_halo), _ip(nlp), _fp(neg), _fp(nlv))
    if scaling != 1.0:
        nlv *= scaling
    return Assembled(n=n, rows=rows, cols=cols, ldu_mapping=perm, vals=vals,
                     nl_rows=nlr[:n_halo], nl_cols=nlc[:n_halo], nl_mapping=nlp[:n_halo],
                     nl_vals=nlv[:n_halo], target_ids=tid[:nt.value], target_sizes=tsz[:nt.value],
                     send_idxs=sidx[:n_halo], b=_f64(s.source) * scaling, x=_f64(s.psi).copy())


# ------------------------------------------------------------------ solve

@dataclass
class OracleSolve:
    x: List[np.ndarray]
    init_residual: float
    final_residual: float
    criterion_calls: int
    n_iterations: int
    norm_factor: float
    history: np.ndarray
    seconds: float


def _rank_structs(asms: Sequence[Assembled], keep):
    arr = (RankSystem * len(asms))()
    for r, a in enumerate(asms):
        fields = dict(rows=_i32(a.rows), cols=_i32(a.cols), vals=_f64(a.vals),
                      nl_rows=_i32(a.nl_rows), nl_cols=_i32(a.nl_cols), nl_vals=_f64(a.nl_vals),
                      target_ids=_i32(a.target_ids), target_sizes=_i32(a.target_sizes),
                      send_idxs=_i32(a.send_idxs), b=_f64(a.b), x=_f64(a.x).copy())
        keep.append(fields)
        arr[r].n, arr[r].nnz, arr[r].n_halo = a.n, fields["rows"].size, fields["nl_rows"].size
        arr[r].n_targets = fields["target_ids"].size
        for k in ("rows", "cols", "nl_rows", "nl_cols", "target_ids", "target_sizes", "send_idxs"):
            setattr(arr[r], k, _ip(fields[k]))
        for k in ("vals", "nl_vals", "b", "x"):
            setattr(arr[r], k, _fp(fields[k]))
    return arr


def solve(asms: Sequence[Assembled], solver="GKOCG", preconditioner="BJ", max_block_size=1,
          tolerance=1e-6, rel_tol=0.0, min_iter=0, max_iter=1000, frequency=1, krylov_dim=100,
          threads=1, mg_max_levels=9, mg_min_coarse_rows=10, mg_coarse_iters=4) -> OracleSolve:
    keep: list = []
    arr = _rank_structs(asms, keep)
    mi = max_iter * 2 if solver == "GKOBiCGStab" else max_iter   # StoppingCriterion.H:188
    p = SolveParams(SOLVERS[solver], PRECONDS[preconditioner], max_block_size, tolerance, rel_tol,
                    min_iter, mi, frequency, krylov_dim, threads, mg_max_levels, mg_min_coarse_rows,
                    mg_coarse_iters)
    res = SolveResult()
    cap = mi + 8
    hist = np.zeros(cap)
    rc = lib().orc_solve(C.c_int(len(asms)), arr, C.byref(p), C.byref(res), _fp(hist),
                         C.c_int32(cap))
    if rc != 0:
        raise RuntimeError(f"orc_solve failed with code {rc}")
    return OracleSolve(x=[k["x"] for k in keep], init_residual=res.init_residual,
                       final_residual=res.final_residual, criterion_calls=res.criterion_calls,
                       n_iterations=res.n_iterations, norm_factor=res.norm_factor,
                       history=hist[:res.n_history].copy(), seconds=res.seconds)


def foam_pcg(s, tolerance=1e-6, rel_tol=0.0, min_iter=0, max_iter=1000, solver="PCG") -> OracleSolve:
    """OpenFOAM-native-equivalent baseline (foam_pcg.cpp): face-based Amul + diagonal PCG with
    OpenFOAM's own normFactor / convergence test on one rank's LduSystem (cyclic interfaces only)."""
    if any(i.kind == "processor" for i in s.interfaces):
        raise ValueError("foam_pcg is a single-rank baseline")
    cyc = [k for k, i in enumerate(s.interfaces) if i.kind != "processor"]
    if cyc:
        ir = _i32(np.concatenate([s.interfaces[k].face_cells for k in cyc]))
        ic = _i32(np.concatenate([s.interfaces[s.interfaces[k].nbr_patch].face_cells for k in cyc]))
        ib = _f64(np.concatenate([s.interfaces[k].bou_coeffs for k in cyc]))
    else:
        ir = ic = np.zeros(0, np.int32)
        ib = np.zeros(0)
    la, ua = _i32(s.lower_addr), _i32(s.upper_addr)
    diag, upper = _f64(s.diag), _f64(s.upper)
    lower = None if s.lower is None else _f64(s.lower)
    b, x = _f64(s.source), _f64(s.psi).copy()
    res = SolveResult()
    cap = 2 * max_iter + 8
    hist = np.zeros(cap)
    fn = lib().orc_foam_pcg if solver == "PCG" else lib().orc_foam_pbicgstab
    fn.restype = C.c_int
    rc = fn(C.c_int32(s.n), C.c_int32(la.size), _ip(la), _ip(ua), _fp(diag), _fp(upper),
            _fp(lower) if lower is not None else None, C.c_int32(ir.size), _ip(ir), _ip(ic), _fp(ib),
            _fp(b), _fp(x), C.c_double(tolerance), C.c_double(rel_tol), C.c_int32(min_iter),
            C.c_int32(max_iter), C.byref(res), _fp(hist), C.c_int32(cap))
    if rc != 0:
        raise RuntimeError(f"orc_foam_pcg failed with code {rc}")
    return OracleSolve(x=[x], init_residual=res.init_residual, final_residual=res.final_residual,
                       criterion_calls=res.criterion_calls, n_iterations=res.n_iterations,
                       norm_factor=res.norm_factor, history=hist[:res.n_history].copy(),
                       seconds=res.seconds)


def foam_pbicgstab(s, **kw) -> OracleSolve:
    """OpenFOAM-native-equivalent PBiCGStab + diagonal preconditioner (foam_pcg.cpp)."""
    return foam_pcg(s, solver="PBiCGStab", **kw)


def dist_spmv(asms: Sequence[Assembled], xs: Sequence[np.ndarray]) -> List[np.ndarray]:
    keep: list = []
    arr = _rank_structs(asms, keep)
    xs = [_f64(x) for x in xs]
    ys = [np.zeros(a.n) for a in asms]
    xp = (f64p * len(asms))(*[_fp(x) for x in xs])
    yp = (f64p * len(asms))(*[_fp(y) for y in ys])
    lib().orc_dist_spmv(C.c_int(len(asms)), arr, xp, yp)
    return ys


def isai_values(n, row_ptrs, cols, vals, spd: bool):
    """ISAI / GISAI values over the CSR pattern: (W, WT) (WT only meaningful for spd)."""
    rp, c, v = _i32(row_ptrs), _i32(cols), _f64(vals)
    w, wt = np.zeros(v.size), np.zeros(v.size)
    fn = lib().orc_isai_generate
    fn.restype = C.c_int
    rc = fn(C.c_int32(n), _ip(rp), _ip(c), _fp(v), C.c_int(1 if spd else 0), _fp(w), _fp(wt))
    if rc != 0:
        raise RuntimeError("ISAI: row too long for the dense solver or missing diagonal")
    return w, wt


class MgHierarchy:
    """The oracle's Multigrid hierarchy of one (local) CSR matrix (multigrid.hpp), level by level."""

    def __init__(self, n, row_ptrs, cols, vals, max_levels=9, min_coarse_rows=10, coarse_iters=4):
        L = lib()
        L.orc_mg_create.restype = C.c_void_p
        rp, c, v = _i32(row_ptrs), _i32(cols), _f64(vals)
        self.h = C.c_void_p(L.orc_mg_create(C.c_int32(n), _ip(rp), _ip(c), _fp(v), C.c_int(max_levels),
                                            C.c_int32(min_coarse_rows), C.c_int(coarse_iters)))
        self.n = n
        self.levels = []
        for l in range(L.orc_mg_levels(self.h)):
            ln, lnnz, lnc = C.c_int32(0), C.c_int32(0), C.c_int32(0)
            L.orc_mg_level_info(self.h, C.c_int(l), C.byref(ln), C.byref(lnnz), C.byref(lnc))
            lrp, lc = np.zeros(ln.value + 1, np.int32), np.zeros(lnnz.value, np.int32)
            lv = np.zeros(lnnz.value)
            agg = np.zeros(ln.value, np.int32) if lnc.value > 0 else None
            L.orc_mg_level_get(self.h, C.c_int(l), _ip(lrp), _ip(lc), _fp(lv), _ip(agg) if agg is not None else None)
            self.levels.append(dict(n=ln.value, nnz=lnnz.value, n_coarse=lnc.value, row_ptrs=lrp, cols=lc,
                                    vals=lv, agg=agg))

    def apply(self, r):
        r = _f64(r)
        z = np.zeros(self.n)
        lib().orc_mg_apply(self.h, _fp(r), _fp(z))
        return z

    def close(self):
        if self.h:
            lib().orc_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def trifactor(kind, n, row_ptrs, cols, vals):
    """ILU(0) / IC(0) factors over the CSR pattern (kind: "ILU", "IC", "IRILU"), trifactor.hpp."""
    rp, c, v = _i32(row_ptrs), _i32(cols), _f64(vals)
    out = np.zeros(v.size)
    fn = lib().orc_trifactor
    fn.restype = C.c_int
    rc = fn(C.c_int(PRECONDS[kind]), C.c_int32(n), _ip(rp), _ip(c), _fp(v), _fp(out))
    if rc != 0:
        raise RuntimeError(f"{kind}: factorisation failed (code {rc}: 1 missing diagonal / repeated column, "
                           "2 unsymmetric pattern)")
    return out


def trifactor_apply(kind, n, row_ptrs, cols, factors, r):
    """z = M^-1 r with the factors of `trifactor` (exact triangular solves; IRILU: 5 + 5 sweeps)."""
    rp, c, f, r = _i32(row_ptrs), _i32(cols), _f64(factors), _f64(r)
    z = np.zeros(n)
    fn = lib().orc_trifactor_apply
    fn.restype = C.c_int
    rc = fn(C.c_int(PRECONDS[kind]), C.c_int32(n), _ip(rp), _ip(c), _fp(f), _fp(r), _fp(z))
    if rc != 0:
        raise RuntimeError(f"{kind}: apply failed (code {rc})")
    return z


def bj_blocks(n, row_ptrs, cols, vals, max_block_size):
    row_ptrs, cols, vals = _i32(row_ptrs), _i32(cols), _f64(vals)
    bp = np.zeros(n + 1, np.int32)
    nb = lib().orc_bj_find_blocks(C.c_int32(n), _ip(row_ptrs), _ip(cols),
                                  C.c_int32(max_block_size), _ip(bp))
    bp = bp[:nb + 1].copy()
    sizes = np.diff(bp)
    inv = np.zeros(int((sizes.astype(np.int64) ** 2).sum()))
    lib().orc_bj_invert_blocks(C.c_int32(n), _ip(row_ptrs), _ip(cols), _fp(vals), C.c_int32(nb),
                               _ip(bp), _fp(inv))
    return bp, inv


def time_spmv(n, row_ptrs, cols, vals, x, reps, threads):
    row_ptrs, cols, vals, x = _i32(row_ptrs), _i32(cols), _f64(vals), _f64(x)
    y = np.zeros(n)
    t = lib().orc_time_spmv(C.c_int32(n), _ip(row_ptrs), _ip(cols), _fp(vals), _fp(x), _fp(y),
                            C.c_int(reps), C.c_int(threads))
    return t, y
