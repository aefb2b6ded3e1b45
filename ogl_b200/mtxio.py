"""Matrix-Market import of an exported linear system.

The reference dumps the assembled system of a field with the `export`/`debug`
keywords (common.C:31-58 `export_system`, CsrMatrixWrapper.H:273-290,
Vector.H:173-176): `<field>_A_local.mtx`, `<field>_A_non_local.mtx` (coordinate
layout, 1-based, `setprecision(15)`) and `<field>_rhs_b_.mtx` (array layout)
under `processorN/<time>/`.  `ogl_export_mtx` writes the same three files; this
module is the way back: it turns such a dump -- from this library or from a real
OGL/OpenFOAM run -- into the `LduSystem` the plugin surface consumes, so that
exported systems can be solved again for parity studies.

Only host logic (numpy); nothing here runs on the solve path."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

import json

from .cases import Interface, LduSystem
from .host import FatalError

SIDECAR = "_partition.json"


def read_mtx(path: str):
    """-> ("coordinate", n_rows, n_cols, rows, cols, vals) with 0-based int32 indices,
    or ("array", n_rows, n_cols, dense[n_rows, n_cols])."""
    with open(path) as f:
        header = f.readline().split()
        if len(header) < 5 or header[0] != "%%MatrixMarket" or header[1] != "matrix":
            raise FatalError(f"{path}: not a Matrix-Market file")
        layout, field, symmetry = header[2], header[3], header[4]
        if field not in ("real", "integer") or symmetry != "general":
            raise FatalError(f"{path}: only real general matrices are written by OGL")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        dims = [int(t) for t in line.split()]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # empty body (nnz = 0)
            body = np.loadtxt(f, dtype=np.float64, ndmin=2)
    if layout == "coordinate":
        n_rows, n_cols, nnz = dims
        if nnz == 0:
            z = np.zeros(0, dtype=np.int32)
            return "coordinate", n_rows, n_cols, z, z.copy(), np.zeros(0)
        if body.shape != (nnz, 3):
            raise FatalError(f"{path}: expected {nnz} coordinate entries, found {body.shape[0]}")
        rows = body[:, 0].astype(np.int64) - 1
        cols = body[:, 1].astype(np.int64) - 1
        if rows.min() < 0 or rows.max() >= n_rows or cols.min() < 0 or cols.max() >= n_cols:
            raise FatalError(f"{path}: index out of range")
        return "coordinate", n_rows, n_cols, rows.astype(np.int32), cols.astype(np.int32), body[:, 2].copy()
    if layout == "array":
        n_rows, n_cols = dims
        vals = body.reshape(-1)
        if vals.size != n_rows * n_cols:
            raise FatalError(f"{path}: expected {n_rows * n_cols} array entries, found {vals.size}")
        return "array", n_rows, n_cols, vals.reshape(n_cols, n_rows).T.copy()   # column-major file order
    raise FatalError(f"{path}: unknown layout {layout}")


def write_mtx_coordinate(path: str, n_rows: int, n_cols: int, rows, cols, vals) -> None:
    """Same text as `gko::write` / `ogl_export_mtx`: 1-based, given order, 15 digits."""
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{n_rows} {n_cols} {len(vals)}\n")
        for r, c, v in zip(rows, cols, vals):
            f.write(f"{int(r) + 1} {int(c) + 1} {float(v):.15g}\n")


def write_mtx_array(path: str, vec) -> None:
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write(f"{len(vec)} 1\n")
        for v in vec:
            f.write(f"{float(v):.15g}\n")


def ldu_from_coo(n: int, rows, cols, vals) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """Square coordinate matrix with a full diagonal and a structurally symmetric
    pattern (every lduMatrix is one) -> (lowerAddr, upperAddr, diag, upper, lower|None).
    Faces come out in OpenFOAM's upper-triangular order (owner, then neighbour,
    ascending), which is the order `init_local_sparsity` expects
    (HostMatrixFreeFunctions.C:105-201); `lower` is None when the values are symmetric."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    key = rows * n + cols
    if np.unique(key).size != key.size:
        raise FatalError("duplicate (row, col) entries: not an assembled lduMatrix")
    d = rows == cols
    if int(d.sum()) != n:
        raise FatalError("the diagonal is not full: not an lduMatrix")
    diag = np.empty(n)
    diag[rows[d]] = vals[d]
    up = rows < cols
    lo = rows > cols
    if int(up.sum()) != int(lo.sum()):
        raise FatalError("pattern is not structurally symmetric: not an lduMatrix")
    order_u = np.lexsort((cols[up], rows[up]))
    lower_addr = rows[up][order_u].astype(np.int32)
    upper_addr = cols[up][order_u].astype(np.int32)
    upper = vals[up][order_u]
    # the transposed positions, in the same face order
    order_l = np.lexsort((rows[lo], cols[lo]))
    if not (np.array_equal(cols[lo][order_l], lower_addr) and np.array_equal(rows[lo][order_l], upper_addr)):
        raise FatalError("pattern is not structurally symmetric: not an lduMatrix")
    lower = vals[lo][order_l]
    return lower_addr, upper_addr, diag, upper, (None if np.array_equal(lower, upper) else lower)


def import_system(folder: str, field: str, psi: Optional[np.ndarray] = None) -> LduSystem:
    """`<folder>/<field>_A_local.mtx` (+ `_rhs_b_.mtx`, + an empty or absent
    `_A_non_local.mtx`) -> LduSystem on one rank.  A dump with non-local entries
    belongs to a decomposed case whose neighbour lists are not part of the files
    (the reference does not export them either): rejected with a FatalError."""
    kind, n, n_cols, rows, cols, vals = read_mtx(os.path.join(folder, field + "_A_local.mtx"))
    if kind != "coordinate" or n != n_cols:
        raise FatalError("local matrix must be a square coordinate matrix")
    nl = os.path.join(folder, field + "_A_non_local.mtx")
    if os.path.exists(nl):
        nk = read_mtx(nl)
        if nk[0] != "coordinate" or len(nk[5]) > 0:
            raise FatalError("non-local entries present: import every processor directory through "
                             "the decomposition that produced them (interfaces are not exported)")
    lower_addr, upper_addr, diag, upper, lower = ldu_from_coo(n, rows, cols, vals)
    rhs = os.path.join(folder, field + "_rhs_b_.mtx")
    if os.path.exists(rhs):
        bk = read_mtx(rhs)
        if bk[0] != "array" or bk[1] != n:
            raise FatalError("right-hand side must be an n x 1 array")
        source = bk[3][:, 0].copy()
    else:
        source = np.zeros(n)
    return LduSystem(n=n, lower_addr=lower_addr, upper_addr=upper_addr, diag=diag, upper=upper, lower=lower,
                     interfaces=[], source=source, psi=np.zeros(n) if psi is None else np.asarray(psi, float).copy(),
                     global_ids=np.arange(n, dtype=np.int64))


# ---- decomposed dumps ---------------------------------------------------------------------
# The reference writes one set of files per processor directory (common.C:31-58: the path is
# `<case>/processorN/<time>/`); the matrices alone do not say which neighbour a halo column belongs
# to.  This library's export therefore adds a small side-car next to them,
# `<field>_partition.json` = {rank, n_ranks, n_local, target_ids, target_sizes} -- OGL's
# communication pattern (HostMatrix.C:251-306) --, and `import_decomposed` rebuilds every rank's
# processor interfaces from it: halo column k of `_A_non_local.mtx` is the k-th processor face
# (HostMatrix.C:438-466), its row the face cell, its value minus the boundary coefficient
# (:199-204), and the columns are blocked by ascending neighbour rank.

def write_partition_sidecar(folder: str, field: str, rank: int, n_ranks: int, n_local: int,
                            target_ids, target_sizes) -> None:
    with open(os.path.join(folder, field + SIDECAR), "w") as f:
        json.dump({"rank": int(rank), "n_ranks": int(n_ranks), "n_local": int(n_local),
                   "target_ids": [int(t) for t in target_ids],
                   "target_sizes": [int(t) for t in target_sizes]}, f)


def import_rank(folder: str, field: str) -> LduSystem:
    """One processor directory of a decomposed dump -> that rank's LduSystem (interfaces included)."""
    side = os.path.join(folder, field + SIDECAR)
    if not os.path.exists(side):
        raise FatalError(f"{side} is missing: a decomposed dump needs the partition side-car this "
                         "library's export writes (the matrices do not carry the neighbour lists)")
    part = json.load(open(side))
    kind, n, n_cols, rows, cols, vals = read_mtx(os.path.join(folder, field + "_A_local.mtx"))
    if kind != "coordinate" or n != n_cols or n != part["n_local"]:
        raise FatalError("local matrix must be a square coordinate matrix of the side-car's size")
    lower_addr, upper_addr, diag, upper, lower = ldu_from_coo(n, rows, cols, vals)
    nk = read_mtx(os.path.join(folder, field + "_A_non_local.mtx"))
    if nk[0] != "coordinate" or nk[1] != n:
        raise FatalError("non-local matrix must be an n x n_halo coordinate matrix")
    n_halo, nl_rows, nl_cols, nl_vals = nk[2], nk[3], nk[4], nk[5]
    sizes = [int(t) for t in part["target_sizes"]]
    if sum(sizes) != n_halo or len(nl_vals) != n_halo or len(sizes) != len(part["target_ids"]):
        raise FatalError("side-car and non-local matrix disagree on the halo size")
    if n_halo and not np.array_equal(np.sort(nl_cols), np.arange(n_halo)):
        raise FatalError("every halo column must appear exactly once (one entry per processor face)")
    by_col = np.argsort(nl_cols, kind="stable")       # running interface index order
    face_cells, bou = nl_rows[by_col], -nl_vals[by_col]
    interfaces, off = [], 0
    for nbr, size in zip(part["target_ids"], sizes):
        interfaces.append(Interface("processor", face_cells[off:off + size].astype(np.int32),
                                    bou[off:off + size].copy(), nbr_rank=int(nbr)))
        off += size
    rhs = os.path.join(folder, field + "_rhs_b_.mtx")
    source = read_mtx(rhs)[3][:, 0].copy() if os.path.exists(rhs) else np.zeros(n)
    if source.size != n:
        raise FatalError("right-hand side must be an n x 1 array")
    return LduSystem(n=n, lower_addr=lower_addr, upper_addr=upper_addr, diag=diag, upper=upper, lower=lower,
                     interfaces=interfaces, source=source, psi=np.zeros(n),
                     global_ids=np.arange(n, dtype=np.int64), rank=int(part["rank"]),
                     n_ranks=int(part["n_ranks"]))


def import_decomposed(case_folder: str, time_name: str, field: str):
    """`<case>/processor0..R-1/<time>/<field>_*` -> [LduSystem per rank] with consistent
    global ids (rank offsets by exclusive scan of the local sizes, gkoGlobalIndex.C:172-201)."""
    systems, r = [], 0
    while os.path.isdir(os.path.join(case_folder, f"processor{r}", time_name)):
        systems.append(import_rank(os.path.join(case_folder, f"processor{r}", time_name), field))
        r += 1
    if not systems:
        raise FatalError(f"no processor*/{time_name} directories under {case_folder}")
    if any(s.n_ranks != len(systems) or s.rank != i for i, s in enumerate(systems)):
        raise FatalError("processor directories and side-cars disagree on the decomposition")
    # the neighbour relation must be symmetric, block sizes included
    for s in systems:
        for itf in s.interfaces:
            back = [j for j in systems[itf.nbr_rank].interfaces if j.nbr_rank == s.rank]
            if sum(j.face_cells.size for j in back) != sum(
                    j.face_cells.size for j in s.interfaces if j.nbr_rank == itf.nbr_rank):
                raise FatalError(f"ranks {s.rank} and {itf.nbr_rank} disagree on their shared faces")
    off = 0
    for s in systems:
        s.global_ids = off + np.arange(s.n, dtype=np.int64)
        off += s.n
    return systems
