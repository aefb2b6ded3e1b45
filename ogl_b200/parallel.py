"""Rank bootstrap: one process per GPU, `torch.distributed` for the rendezvous
(the role MPI_COMM_WORLD / Pstream play in the reference,
DevicePersistent/ExecutorHandler/ExecutorHandler.H:29-33,140-144), NCCL inside
libogl_b200.so for the data path."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional


@dataclass
class Pstream:
    """Pstream::parRun()/myProcNo()/nProcs() analogue."""
    rank: int = 0
    n_ranks: int = 1
    local_rank: int = 0
    nccl_id: Optional[bytes] = None   # None on several ranks: host-driven window bootstrap (below)

    @property
    def par_run(self) -> bool:
        return self.n_ranks > 1

    @property
    def master(self) -> bool:
        return self.rank == 0

    # the two collectives the host layer needs when the library has no communicator of its own
    # (Pstream::gatherList / Pstream::barrier in an OpenFOAM build)
    def all_gather_bytes(self, blob: bytes):
        import torch.distributed as dist
        out = [None] * self.n_ranks
        dist.all_gather_object(out, blob)
        return out

    def barrier(self):
        import torch.distributed as dist
        dist.barrier()

    def broadcast_scalar(self, value: float, src: int = 0) -> float:
        import torch.distributed as dist
        box = [float(value)]
        dist.broadcast_object_list(box, src=src)
        return float(box[0])


def init_from_env(backend: Optional[str] = None, nccl: bool = True) -> Pstream:
    """Join the job described by RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* (torchrun)
    and agree on an NCCL unique id for the solver library.  nccl=False: no id -- the library's
    peer-memory windows are bootstrapped through this process group instead (several ranks may
    then share one device; set OGL_B200_DEVICE to pin them)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world == 1:
        return Pstream(0, 1, local, None)
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    if not nccl:
        return Pstream(rank, world, int(os.environ.get("OGL_B200_DEVICE", local)), None)
    return Pstream(rank, world, local, broadcast_nccl_id(rank))


def broadcast_nccl_id(rank: int) -> bytes:
    """Rank 0 creates the id (ogl_nccl_unique_id), everyone receives it."""
    import torch
    import torch.distributed as dist

    from . import _lib as L

    n = L.OGL_NCCL_ID_BYTES
    if rank == 0:
        from .backend import nccl_unique_id
        payload = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8)
    else:
        payload = torch.zeros(n, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        payload = payload.cuda()
    dist.broadcast(payload, src=0)
    return bytes(payload.cpu().tolist())
