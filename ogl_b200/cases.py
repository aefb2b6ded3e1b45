"""Synthetic blockMesh-like LDU systems.

OpenFOAM is not available in this image, so these generators ARE the
"blockMesh" cases of BASELINE.json: structured hex boxes in OpenFOAM's own
face/addressing conventions, decomposed like ``decomposePar -method simple``.

Conventions (all [upstream OpenFOAM], restated in SURVEY.md section 8d/8e):

* cell id ``c = i + Nx*(j + Ny*k)``; internal faces are enumerated cell by cell
  in the order (i+1, j+1, k+1), which yields ``lowerAddr`` ascending and, per
  owner, ``upperAddr`` ascending (upper-triangular order);
* ``simple`` decomposition [px,py,pz]: contiguous index boxes, rank id
  ``rx + px*(ry + py*rz)``, local cells renumbered i-fastest inside the box;
* processor patches are ordered by ascending neighbour rank, a cyclic pair cut
  by the decomposition becomes a second (processorCyclic) patch to the same
  neighbour, listed after the plain one; faces of a patch are ordered by the
  global face order, which both sides see identically;
* the matrix entry of a coupled face is ``-interfaceBouCoeffs``
  (HostMatrix/HostMatrix.C:199-204 in the reference).

Everything here is plain numpy; nothing imports the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class Interface:
    """One coupled patch of a rank (lduInterfaceField analogue)."""

    kind: str                    # "processor" | "cyclic"
    face_cells: np.ndarray       # int32, local cell ids (faceCells())
    bou_coeffs: np.ndarray       # float64, interfaceBouCoeffs[i]
    nbr_rank: int = -1           # processor: neighbProcNo()
    nbr_patch: int = -1          # cyclic: neighbPatchID() (index into interfaces)


@dataclass
class LduSystem:
    """One rank's lduMatrix + interfaces + rhs (what lduMatrix::solver::New gets)."""

    n: int
    lower_addr: np.ndarray       # int32 [F] owner (row of an upper entry)
    upper_addr: np.ndarray       # int32 [F] neighbour
    diag: np.ndarray             # float64 [n]
    upper: np.ndarray            # float64 [F]
    lower: Optional[np.ndarray]  # float64 [F] or None when symmetric
    interfaces: List[Interface]
    source: np.ndarray           # b
    psi: np.ndarray              # initial guess
    global_ids: np.ndarray       # int64 [n] global cell id of each local cell
    rank: int = 0
    n_ranks: int = 1
    x_star: Optional[np.ndarray] = None   # manufactured solution (local part)

    @property
    def symmetric(self) -> bool:
        return self.lower is None

    @property
    def n_faces(self) -> int:
        return int(self.lower_addr.size)


# --------------------------------------------------------------------------
# addressing
# --------------------------------------------------------------------------

def box_addressing(nx: int, ny: int, nz: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """lowerAddr, upperAddr (int32) and the direction (0,1,2) of every internal
    face of an nx*ny*nz box in OpenFOAM's upper-triangular order."""
    n = nx * ny * nz
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx * ny)
    valid = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], axis=1)
    nbr = c[:, None] + np.array([1, nx, nx * ny], dtype=np.int64)[None, :]
    direction = np.broadcast_to(np.arange(3, dtype=np.int8)[None, :], valid.shape)
    lower = np.broadcast_to(c[:, None], valid.shape)[valid]
    upper = nbr[valid]
    return lower.astype(np.int32), upper.astype(np.int32), direction[valid].copy()


def _split(nglob: int, parts: int) -> List[Tuple[int, int]]:
    base, rem = divmod(nglob, parts)
    out, lo = [], 0
    for p in range(parts):
        hi = lo + base + (1 if p < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass
class Box:
    """Index box of one rank inside the global mesh."""

    lo: Tuple[int, int, int]
    hi: Tuple[int, int, int]

    @property
    def shape(self) -> Tuple[int, int, int]:
        return tuple(h - l for l, h in zip(self.lo, self.hi))  # type: ignore

    @property
    def n(self) -> int:
        s = self.shape
        return s[0] * s[1] * s[2]


def simple_decomposition(dims: Sequence[int], procs: Sequence[int]) -> List[Box]:
    xs, ys, zs = (_split(d, p) for d, p in zip(dims, procs))
    boxes = []
    for rz in range(procs[2]):
        for ry in range(procs[1]):
            for rx in range(procs[0]):
                boxes.append(Box((xs[rx][0], ys[ry][0], zs[rz][0]),
                                 (xs[rx][1], ys[ry][1], zs[rz][1])))
    return boxes


def _rank_of(coord, procs):
    return coord[0] + procs[0] * (coord[1] + procs[1] * coord[2])


def _rank_coord(rank, procs):
    return (rank % procs[0], (rank // procs[0]) % procs[1], rank // (procs[0] * procs[1]))


# --------------------------------------------------------------------------
# coefficient models
# --------------------------------------------------------------------------

class PressureModel:
    """Symmetric 7-point pressure matrix.

    off-diagonal ``+coef`` on every face, ``diag = -sum(off) - g_c`` with a
    Dirichlet-like ``g_c = 2*coef`` on the k = Nz-1 layer (the lid) -- the
    negative-diagonal sign convention of README.md:101; ``sign=-1`` gives the
    SPD twin (diag > 0, off-diag < 0; test/data_validation.py:100-111)."""

    symmetric = True

    def __init__(self, dims, coef=1e-5, sign=1.0, cyclic=(False, False, False),
                 ref_cell=False):
        self.dims = tuple(dims)
        self.coef = coef
        self.sign = sign
        self.cyclic = tuple(cyclic)
        self.ref_cell = ref_cell

    def face_coeffs(self, owner_g, nbr_g, direction):
        v = np.full(owner_g.shape, self.sign * self.coef)
        return v, v  # upper, lower

    def diag(self, gid, n_nbrs):
        nx, ny, nz = self.dims
        k = gid // (nx * ny)
        g = np.where(k == nz - 1, 2.0 * self.coef, 0.0)
        if self.ref_cell:
            g = g + np.where(gid == 0, self.coef, 0.0)
        return self.sign * (-(n_nbrs * self.coef) - g)


class MomentumModel:
    """Asymmetric convection-diffusion matrix:
    ``upper_f = -nu - max(-phi_f,0)``, ``lower_f = -nu - max(phi_f,0)``,
    ``diag = V/dt + sum(...)`` strictly diagonally dominant; phi from a smooth
    analytic velocity field evaluated at the face centre."""

    symmetric = False

    def __init__(self, dims, nu=1.0, vdt=0.5, umag=2.0, cyclic=(False, False, False)):
        self.dims = tuple(dims)
        self.nu = nu
        self.vdt = vdt
        self.umag = umag
        self.cyclic = tuple(cyclic)

    def _phi(self, owner_g, direction):
        nx, ny, nz = self.dims
        i = owner_g % nx
        j = (owner_g // nx) % ny
        k = owner_g // (nx * ny)
        x = (i + 0.5 + 0.5 * (direction == 0)) / nx
        y = (j + 0.5 + 0.5 * (direction == 1)) / ny
        z = (k + 0.5 + 0.5 * (direction == 2)) / nz
        two_pi = 2.0 * np.pi
        u = np.sin(two_pi * x) * np.cos(two_pi * y) * np.cos(two_pi * z)
        v = -np.cos(two_pi * x) * np.sin(two_pi * y) * np.cos(two_pi * z)
        w = 0.3 * np.cos(two_pi * x) * np.cos(two_pi * y) * np.sin(two_pi * z)
        return self.umag * np.where(direction == 0, u, np.where(direction == 1, v, w))

    def face_coeffs(self, owner_g, nbr_g, direction):
        phi = self._phi(owner_g, direction)
        upper = -self.nu - np.maximum(-phi, 0.0)
        lower = -self.nu - np.maximum(phi, 0.0)
        return upper, lower

    def diag_from_offdiag(self, neg_offdiag_sum):
        return self.vdt + neg_offdiag_sum


# --------------------------------------------------------------------------
# system builder
# --------------------------------------------------------------------------

def _local_ids(box: Box, I, J, K):
    sx, sy, _ = box.shape
    return ((I - box.lo[0]) + sx * ((J - box.lo[1]) + sy * (K - box.lo[2]))).astype(np.int64)


def _global_ids(dims, I, J, K):
    return (I + dims[0] * (J + dims[1] * K)).astype(np.int64)


def _box_cells(box: Box):
    sx, sy, sz = box.shape
    l = np.arange(sx * sy * sz, dtype=np.int64)
    I = box.lo[0] + l % sx
    J = box.lo[1] + (l // sx) % sy
    K = box.lo[2] + l // (sx * sy)
    return I, J, K


def _plane(box: Box, axis: int, at_hi: bool):
    """(I,J,K) of the box cells on its low/high face along `axis`, ordered by
    local id (== the global face order of that patch, see module docstring)."""
    I, J, K = _box_cells(box)
    coord = (I, J, K)[axis]
    sel = coord == (box.hi[axis] - 1 if at_hi else box.lo[axis])
    return I[sel], J[sel], K[sel]


def build_rank_system(model, dims, procs, rank, x_star_global=None, seed=20240621,
                      rhs_scale=1.0) -> LduSystem:
    """Assemble the LduSystem rank `rank` would see for `model` on the global
    mesh `dims` decomposed as `procs`."""
    dims = tuple(int(d) for d in dims)
    procs = tuple(int(p) for p in procs)
    n_ranks = procs[0] * procs[1] * procs[2]
    boxes = simple_decomposition(dims, procs)
    box = boxes[rank]
    sx, sy, sz = box.shape
    n = box.n
    rc = _rank_coord(rank, procs)

    lower_addr, upper_addr, direction = box_addressing(sx, sy, sz)
    I, J, K = _box_cells(box)
    gid = _global_ids(dims, I, J, K)
    owner_g = gid[lower_addr]
    nbr_g = gid[upper_addr]
    up_v, lo_v = model.face_coeffs(owner_g, nbr_g, direction.astype(np.int64))

    # -sum(offdiag) per row and neighbour count, accumulated over every face
    # (internal, processor, cyclic) of the GLOBAL matrix row
    neg_off = np.zeros(n)
    n_nbrs = np.zeros(n)
    np.add.at(neg_off, lower_addr, -up_v)
    np.add.at(neg_off, upper_addr, -lo_v)
    np.add.at(n_nbrs, lower_addr, 1.0)
    np.add.at(n_nbrs, upper_addr, 1.0)

    # coupled patches: (nbr_rank, is_cyclic_cut, axis, at_hi)
    proc_patches = []   # (nbr_rank, order_key, face_cells, bou, halo global ids)
    cyc_patches = []    # local cyclic: (axis, at_hi, face_cells, nbr face_cells, bou)
    for axis in range(3):
        for at_hi in (False, True):
            on_domain_bnd = (box.hi[axis] == dims[axis]) if at_hi else (box.lo[axis] == 0)
            if on_domain_bnd and not model.cyclic[axis]:
                continue
            if dims[axis] == 1:
                continue
            pI, pJ, pK = _plane(box, axis, at_hi)
            fc = _local_ids(box, pI, pJ, pK)
            g_mine = _global_ids(dims, pI, pJ, pK)
            nI, nJ, nK = pI.copy(), pJ.copy(), pK.copy()
            ncoord = (nI, nJ, nK)[axis]
            ncoord += 1 if at_hi else -1
            ncoord %= dims[axis]
            g_nbr = _global_ids(dims, nI, nJ, nK)
            # owner of the face = the side whose cell precedes along +axis
            # (for a cyclic wrap the owner is the high-side cell's partner:
            # keep "low index side owns" for interior cuts, and for the wrap the
            # cell at hi owns the face whose neighbour is the cell at 0)
            if at_hi:
                o_g, dirn = g_mine, axis
                u, l = model.face_coeffs(o_g, g_nbr, np.full(o_g.shape, dirn, dtype=np.int64))
                mine = u            # A[mine, nbr] is an upper-type entry
            else:
                o_g, dirn = g_nbr, axis
                u, l = model.face_coeffs(o_g, g_mine, np.full(o_g.shape, dirn, dtype=np.int64))
                mine = l            # A[mine, nbr] is a lower-type entry
            np.add.at(neg_off, fc, -mine)
            np.add.at(n_nbrs, fc, 1.0)
            nrc = list(rc)
            nrc[axis] = (rc[axis] + (1 if at_hi else -1)) % procs[axis]
            nbr_rank = _rank_of(nrc, procs)
            if procs[axis] == 1:
                # cyclic pair entirely on this rank -> local interface
                nfc = _local_ids(box, nI, nJ, nK)
                cyc_patches.append((axis, at_hi, fc, nfc, -mine))
            else:
                is_wrap = on_domain_bnd
                proc_patches.append((nbr_rank, 1 if is_wrap else 0, axis, at_hi,
                                     fc, -mine, g_nbr))

    if isinstance(model, MomentumModel):
        diag = model.diag_from_offdiag(neg_off)
    else:
        diag = model.diag(gid, n_nbrs)

    interfaces: List[Interface] = []
    halo_gids = []
    # processor patches: ascending neighbour rank, plain before processorCyclic
    for (nbr_rank, wrap, axis, at_hi, fc, bou, g_nbr) in sorted(
            proc_patches, key=lambda t: (t[0], t[1], t[2], t[3])):
        interfaces.append(Interface("processor", fc.astype(np.int32), bou.astype(np.float64),
                                    nbr_rank=int(nbr_rank)))
        halo_gids.append(g_nbr)
    # cyclic patches after the processor ones, pairs adjacent (lo, hi)
    base = len(interfaces)
    for idx, (axis, at_hi, fc, nfc, bou) in enumerate(cyc_patches):
        partner = base + (idx + 1 if not at_hi else idx - 1)
        interfaces.append(Interface("cyclic", fc.astype(np.int32), bou.astype(np.float64),
                                    nbr_patch=partner))

    # manufactured solution and rhs b = A x*
    n_global = dims[0] * dims[1] * dims[2]
    if x_star_global is None:
        x_star_global = np.random.default_rng(seed).uniform(-1.0, 1.0, n_global)
    xs = x_star_global[gid]
    b = diag * xs
    b += np.bincount(lower_addr, weights=up_v * xs[upper_addr], minlength=n)
    b += np.bincount(upper_addr, weights=lo_v * xs[lower_addr], minlength=n)
    pi = 0
    for itf in interfaces:
        if itf.kind == "processor":
            b += np.bincount(itf.face_cells, weights=-itf.bou_coeffs * x_star_global[halo_gids[pi]],
                             minlength=n)
            pi += 1
        else:
            nfc = interfaces[itf.nbr_patch].face_cells
            b += np.bincount(itf.face_cells, weights=-itf.bou_coeffs * xs[nfc], minlength=n)

    return LduSystem(
        n=n, lower_addr=lower_addr, upper_addr=upper_addr, diag=diag,
        upper=np.ascontiguousarray(up_v, dtype=np.float64),
        lower=None if model.symmetric else np.ascontiguousarray(lo_v, dtype=np.float64),
        interfaces=interfaces, source=rhs_scale * b, psi=np.zeros(n), global_ids=gid,
        rank=rank, n_ranks=n_ranks, x_star=xs)


def build_case(model, dims, procs=(1, 1, 1), seed=20240621) -> List[LduSystem]:
    """All ranks of a decomposed case (small cases / CPU tests)."""
    n_global = int(np.prod(dims))
    xg = np.random.default_rng(seed).uniform(-1.0, 1.0, n_global)
    n_ranks = int(np.prod(procs))
    return [build_rank_system(model, dims, procs, r, x_star_global=xg) for r in range(n_ranks)]


# --------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------

def cavity_2d(procs=(2, 1, 1), sign=1.0) -> List[LduSystem]:
    """configs[0]: icoFoam cavity 20x20x1 pressure system, rAU*|Sf|/delta = 5e-5,
    all-Neumann with a reference cell."""
    dims = (20, 20, 1)
    m = PressureModel(dims, coef=5e-5, sign=sign, ref_cell=True)
    # a 2-D cavity has no lid Dirichlet on p: drop the k-layer term
    m.diag = lambda gid, n_nbrs, _m=m: _m.sign * (
        -(n_nbrs * _m.coef) - np.where(gid == 0, _m.coef, 0.0))
    return build_case(m, dims, procs)


def pressure_3d(n: int, procs=(1, 1, 1), rank: Optional[int] = None, sign=1.0):
    """configs[1]/[4]: N^3 lid-driven-cavity pressure system."""
    dims = (n, n, n)
    m = PressureModel(dims, coef=1e-5, sign=sign)
    if rank is None:
        return build_case(m, dims, procs)
    return build_rank_system(m, dims, procs, rank)


def momentum_3d(n: int, procs=(1, 1, 1), rank: Optional[int] = None):
    """configs[2]: N^3 momentum (asymmetric) system, rng seed 20240622."""
    dims = (n, n, n)
    m = MomentumModel(dims)
    if rank is None:
        return build_case(m, dims, procs, seed=20240622)
    return build_rank_system(m, dims, procs, rank, seed=20240622)


def channel(dims=(32, 16, 16), procs=(2, 1, 1)) -> List[LduSystem]:
    """configs[3]: 2:1:1 channel box, cyclic in x and z, walls in y; the
    pressure matrix is made definite by the wall-normal Dirichlet-like term."""
    m = PressureModel(dims, coef=1e-5, cyclic=(True, False, True))
    nx, ny, nz = dims
    m.diag = lambda gid, n_nbrs, _m=m: _m.sign * (
        -(n_nbrs * _m.coef) - np.where(((gid // nx) % ny == 0) | ((gid // nx) % ny == ny - 1),
                                       2.0 * _m.coef, 0.0))
    return build_case(m, dims, procs)


def assemble_global_csr(systems: Sequence[LduSystem]):
    """Global scipy CSR + rhs of a decomposed case, from the ranks' pieces (for
    cross-checks against scipy/numpy; small cases only)."""
    import scipy.sparse as sp

    n_global = sum(s.n for s in systems)
    rows, cols, vals = [], [], []
    b = np.zeros(n_global)
    halo_of = {}
    # what each rank sends to each neighbour, in patch order
    for s in systems:
        for itf in s.interfaces:
            if itf.kind == "processor":
                halo_of.setdefault((s.rank, itf.nbr_rank), []).append(s.global_ids[itf.face_cells])
    cursor = {}
    for s in systems:
        g = s.global_ids
        rows += [g, g[s.lower_addr], g[s.upper_addr]]
        cols += [g, g[s.upper_addr], g[s.lower_addr]]
        vals += [s.diag, s.upper, s.upper if s.lower is None else s.lower]
        b[g] = s.source
        for itf in s.interfaces:
            if itf.kind == "processor":
                key = (itf.nbr_rank, s.rank)
                idx = cursor.get((s.rank, itf.nbr_rank), 0)
                nbr_cells = halo_of[key][idx]
                cursor[(s.rank, itf.nbr_rank)] = idx + 1
                rows.append(g[itf.face_cells]); cols.append(nbr_cells); vals.append(-itf.bou_coeffs)
            else:
                nfc = s.interfaces[itf.nbr_patch].face_cells
                rows.append(g[itf.face_cells]); cols.append(g[nfc]); vals.append(-itf.bou_coeffs)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n_global, n_global)).tocsr()
    return A, b
