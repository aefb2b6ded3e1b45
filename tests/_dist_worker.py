"""Worker of tests/test_dist_gloo.py: one process per rank, gloo backend.

Exercises the N>1 host path without a GPU: rank bootstrap + NCCL-id broadcast
(ogl_b200.parallel), the communication pattern (ogl_b200.host) used for a real
halo exchange over gloo, a distributed SpMV and a distributed Jacobi-PCG in
numpy whose reductions are gloo all-reduces -- compared with the single-process
multi-rank oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from ogl_b200 import cases, host  # noqa: E402
from ogl_b200.parallel import init_from_env  # noqa: E402


def halo_exchange(x, tid, tsz, sidx):
    """recv buffer blocked by ascending neighbour rank (Partition.H:66-67)."""
    recv = np.zeros(int(tsz.sum()))
    off = 0
    reqs, bufs = [], []
    for q, cnt in zip(tid, tsz):
        send = torch.from_numpy(np.ascontiguousarray(x[sidx[off:off + cnt]]))
        buf = torch.zeros(int(cnt), dtype=torch.float64)
        reqs.append(dist.isend(send, int(q)))
        reqs.append(dist.irecv(buf, int(q)))
        bufs.append((off, cnt, buf, send))
        off += cnt
    for r in reqs:
        r.wait()
    for o, cnt, buf, _ in bufs:
        recv[o:o + cnt] = buf.numpy()
    return recv


def allreduce(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t.item())


def main():
    out_path = sys.argv[1]
    case = sys.argv[2]
    ps = init_from_env("gloo")
    rank, world = ps.rank, ps.n_ranks
    procs = (2, 1, 1)
    systems = cases.channel((8, 4, 4), procs) if case == "channel" else cases.pressure_3d(8, procs)
    s = systems[rank]
    a = oracle.assemble(s)                         # checker for the local blocks
    tid, tsz, sidx = host.create_communication_pattern(s)   # product host logic
    fcs = host.collect_cells_on_non_local_interface(s)
    import scipy.sparse as sp
    A = sp.csr_matrix((a.vals, a.cols, a.row_ptrs), shape=(s.n, s.n))
    Anl = sp.csr_matrix((a.nl_vals, (a.nl_rows, a.nl_cols)), shape=(s.n, max(int(tsz.sum()), 1)))

    def spmv(x):
        return A @ x + Anl @ halo_exchange(x, tid, tsz, sidx)

    rng = np.random.default_rng(5)
    xg = rng.normal(size=sum(t.n for t in systems))
    x = xg[s.global_ids]
    y = spmv(x)
    # what the halo contains: the neighbour cells across each processor face
    recv = halo_exchange(s.global_ids.astype(np.float64), tid, tsz, sidx)

    # distributed Jacobi-PCG with the reference's criterion, gloo reductions
    b, xk = s.source.copy(), np.zeros(s.n)
    dinv = 1.0 / A.diagonal()
    r = b - spmv(xk)
    norm_factor = allreduce(np.abs(b).sum() + np.abs(b - r).sum()) + 1e-15   # x0 = 0 -> w = 0
    p = np.zeros(s.n)
    rho_prev, it = 1.0, 0
    while True:
        z = dinv * r
        rho = allreduce(r @ z)
        res = allreduce(np.abs(r).sum()) / norm_factor
        it += 1
        if res < 1e-8 or it > 500:
            break
        p = z + (rho / rho_prev) * p
        q = spmv(p)
        alpha = rho / allreduce(p @ q)
        xk += alpha * p
        r -= alpha * q
        rho_prev = rho

    # the two host collectives of the NCCL-free bootstrap / the adaptive criterion (Pstream helpers)
    blobs = ps.all_gather_bytes(bytes([rank]) * (3 + rank))
    cost = ps.broadcast_scalar(100.0 + rank)       # rank 0's figure everywhere
    ps.barrier()
    # and the decomposed Matrix-Market bridge: every rank dumps its block, rank 0 reads them all back
    from ogl_b200 import mtxio
    folder = os.path.join(os.path.dirname(out_path), f"processor{rank}", "1")
    os.makedirs(folder, exist_ok=True)
    mtxio.write_mtx_coordinate(os.path.join(folder, "p_A_local.mtx"), a.n, a.n, a.rows, a.cols, a.vals)
    mtxio.write_mtx_coordinate(os.path.join(folder, "p_A_non_local.mtx"), a.n, a.nl_rows.size, a.nl_rows, a.nl_cols,
                               a.nl_vals)
    mtxio.write_mtx_array(os.path.join(folder, "p_rhs_b_.mtx"), s.source)
    mtxio.write_partition_sidecar(folder, "p", rank, world, s.n, tid, tsz)
    ps.barrier()
    reread = -1
    if rank == 0:
        back = mtxio.import_decomposed(os.path.dirname(out_path), "1", "p")
        reread = int(sum(len(t.interfaces) for t in back))

    json.dump({"rank": rank, "world": world, "nccl_id": list(ps.nccl_id) if ps.nccl_id else None,
               "blobs": [list(b) for b in blobs], "cost": cost, "reread_interfaces": reread,
               "y": y.tolist(), "halo_gids": recv.tolist(), "iters": it, "res": res,
               "x": xk.tolist(), "fcs": fcs.tolist()},
              open(f"{out_path}.{rank}", "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
