"""Host-side mirror of the reference's HostMatrixWrapper
(HostMatrix/HostMatrix.H:222-440, HostMatrix/HostMatrix.C:16-96) on top of the
C ABI: walks the lduMatrix + interfaces once to describe the sparsity to the
device (which sorts/permutes it there and caches it), and re-uploads only the
coefficient values on later solves."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .backend import Context
from .cases import LduSystem


class FatalError(RuntimeError):
    """FatalErrorInFunction ... abort(FatalError) analogue."""


class ObjectRegistry(dict):
    """objectRegistry analogue: name -> persistent object, owned by the mesh."""

    def found_object(self, name: str) -> bool:
        return name in self

    def time_path(self) -> str:
        return self.get("__time_path__", ".")


def count_interface_nnz(system: LduSystem, proc_interfaces: bool) -> int:
    """HostMatrix.C:158-178."""
    return int(sum(i.face_cells.size for i in system.interfaces
                   if (i.kind == "processor") == proc_interfaces))


def create_communication_pattern(system: LduSystem):
    """HostMatrix.C:251-306: per neighbour rank (ascending) the concatenated
    faceCells of every processor patch to it, in interface order."""
    by_rank: Dict[int, List[np.ndarray]] = {}
    for itf in system.interfaces:
        if itf.kind == "processor":
            by_rank.setdefault(int(itf.nbr_rank), []).append(itf.face_cells)
    ranks = sorted(by_rank)
    target_ids = np.array(ranks, dtype=np.int32)
    send = [np.concatenate(by_rank[r]).astype(np.int32) for r in ranks]
    target_sizes = np.array([s.size for s in send], dtype=np.int32)
    send_idxs = np.concatenate(send) if send else np.zeros(0, np.int32)
    return target_ids, target_sizes, send_idxs


def collect_local_interface_indices(system: LduSystem) -> Tuple[np.ndarray, np.ndarray]:
    """HostMatrix.C:385-410: (row = faceCell, col = patchAddr(neighbPatchID)) of
    every cyclic interface, in interface order; AMI/ACMI are rejected (:339-341)."""
    rows, cols = [], []
    for itf in system.interfaces:
        if itf.kind == "processor":
            continue
        if itf.kind != "cyclic":
            raise FatalError(f"Currently unsupported {itf.kind} patch detected")
        rows.append(itf.face_cells)
        cols.append(system.interfaces[itf.nbr_patch].face_cells)
    if not rows:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    return np.concatenate(rows).astype(np.int32), np.concatenate(cols).astype(np.int32)


def collect_cells_on_non_local_interface(system: LduSystem) -> np.ndarray:
    """HostMatrix.C:412-436: faceCells of the processor interfaces, interface order."""
    fc = [i.face_cells for i in system.interfaces if i.kind == "processor"]
    return np.concatenate(fc).astype(np.int32) if fc else np.zeros(0, np.int32)


def collect_interface_coeffs(system: LduSystem, local: bool) -> np.ndarray:
    """HostMatrix.C:180-207 without the sign flip (the device gather negates)."""
    c = [i.bou_coeffs for i in system.interfaces if (i.kind != "processor") == local]
    return np.concatenate(c).astype(np.float64) if c else np.zeros(0, np.float64)


class HostMatrixWrapper:
    """Per-solve object (the reference constructs one per linear solve); all
    persistent state lives in the registry-held Context."""

    def __init__(self, db: ObjectRegistry, system: LduSystem, controls: dict, field_name: str,
                 ctx: Context, pstream=None):
        self.db, self.system, self.field_name, self.ctx = db, system, field_name, ctx
        self.pstream = pstream
        self.verbose = int(controls.get("verbose", 0))
        self.scaling = float(controls.get("scaling", 1.0))
        if controls.get("reorderOnHost", False):
            raise FatalError("reorderOnHost: this backend permutes on the device only "
                             "(no CPU path); remove the keyword")
        self.nrows = system.n
        self.local_interface_nnz = count_interface_nnz(system, False)
        self.upper_nnz = system.n_faces
        self.local_matrix_nnz = self.nrows + 2 * self.upper_nnz
        self.local_matrix_w_interfaces_nnz = self.local_matrix_nnz + self.local_interface_nnz
        self.non_local_matrix_nnz = count_interface_nnz(system, True)
        key = field_name + "_local_cols"            # HostMatrix.H:28-35
        regenerate = bool(controls.get("regenerate", False))
        if not db.found_object(key) or regenerate:
            self.init_local_sparsity_pattern()
            self.init_non_local_sparsity_pattern()
            db[key] = True
            db[field_name + "_local_coeffs"] = False
        if not db.get(field_name + "_local_coeffs") or controls.get("updateSysMatrix", True):
            self.update_matrix_data()
            db[field_name + "_local_coeffs"] = True

    # HostMatrix.C:468-589 -> ogl_pattern_from_ldu (device sort + cyclic merge)
    def init_local_sparsity_pattern(self):
        s = self.system
        ir, ic = collect_local_interface_indices(s)
        self.ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, s.symmetric, ir, ic)

    # HostMatrix.C:251-306 + Partition.H:57-70 + HostMatrix.C:438-466
    def init_non_local_sparsity_pattern(self):
        tid, tsz, sidx = create_communication_pattern(self.system)
        self.ctx.partition_create(self.system.n, tid, tsz, sidx)
        ps = self.pstream
        if ps is not None and ps.n_ranks > 1 and ps.nccl_id is None:
            # the library has no communicator of its own: move the window directories with the
            # host's (Pstream in an OpenFOAM build), connect, and rendezvous before anyone solves
            self.ctx.partition_connect(ps.all_gather_bytes(self.ctx.partition_export()))
            ps.barrier()
        self.ctx.nonlocal_pattern(collect_cells_on_non_local_interface(self.system))

    # HostMatrix.C:592-732 -> ogl_values_update
    def update_matrix_data(self):
        s = self.system
        self.ctx.values_update(s.diag, s.upper, None if s.symmetric else s.lower,
                               collect_interface_coeffs(s, True),
                               collect_interface_coeffs(s, False), self.scaling)
