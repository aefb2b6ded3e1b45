"""CPU, world_size 2, gloo: the multi-rank host path (rank bootstrap, NCCL-id
broadcast, communication pattern driving a real halo exchange, distributed
reductions) against the single-process multi-rank oracle."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import gather_global
from ogl_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_workers(tmp_path, case):
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_dist_worker.py"), out, case]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [json.load(open(f"{out}.{r}")) for r in range(2)]


@pytest.mark.parametrize("case", ["pressure", "channel"])
def test_two_ranks_over_gloo(oracle, tmp_path, case):
    res = run_workers(tmp_path, case)
    procs = (2, 1, 1)
    systems = cases.channel((8, 4, 4), procs) if case == "channel" else cases.pressure_3d(8, procs)
    asms = [oracle.assemble(s) for s in systems]
    # both ranks hold the same 128-byte NCCL id
    assert res[0]["nccl_id"] is not None and len(res[0]["nccl_id"]) == 128
    assert res[0]["nccl_id"] == res[1]["nccl_id"] and any(res[0]["nccl_id"])
    # host collectives used by the NCCL-free window bootstrap and by the adaptive criterion
    for r in range(2):
        assert res[r]["blobs"] == [[0, 0, 0], [1, 1, 1, 1]] and res[r]["cost"] == 100.0
    # per-processor Matrix-Market dumps written by the ranks and read back on rank 0
    # (one interface per neighbour rank comes back: patches to the same neighbour are concatenated, and
    # rank-local cyclic couplings are ordinary entries of A_local after the dump)
    n_nbr_pairs = sum(len({i.nbr_rank for i in s.interfaces if i.kind == "processor"}) for s in systems)
    assert res[0]["reread_interfaces"] == n_nbr_pairs
    # distributed SpMV == oracle's rank-by-rank emulation
    xg = np.random.default_rng(5).normal(size=sum(s.n for s in systems))
    ys = oracle.dist_spmv(asms, [xg[s.global_ids] for s in systems])
    for r in range(2):
        assert np.allclose(res[r]["y"], ys[r], rtol=1e-13, atol=1e-18)
    # halo slot k holds the neighbour cell of my k-th processor face
    A, _ = cases.assemble_global_csr(systems)
    for r, s in enumerate(systems):
        halo = np.array(res[r]["halo_gids"]).astype(np.int64)
        mine = s.global_ids[np.array(res[r]["fcs"], dtype=np.int64)]
        assert np.all(np.asarray(A[mine, halo]).ravel() != 0)
        assert not np.isin(halo, s.global_ids).any()
    # distributed PCG == oracle
    o = oracle.solve(asms, "GKOCG", "BJ", tolerance=1e-8)
    assert res[0]["iters"] == res[1]["iters"]
    assert abs(res[0]["iters"] - o.n_iterations) <= 2
    x = gather_global(systems, [np.array(res[r]["x"]) for r in range(2)])
    xo = gather_global(systems, o.x)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-8
