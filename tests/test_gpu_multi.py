"""-m gpu, >= 2 GPUs: the decomposed path (NCCL halo exchange overlapped with
the local SpMV, all-reduced fused reductions) through the plugin surface,
against the single-process multi-rank oracle.  Skipped on a 1-GPU box; the
same host logic is covered on CPU by tests/test_dist_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import gather_global

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    from ogl_b200.backend import device_count
    return device_count()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_decomposed_solves_match_oracle(oracle, tmp_path, procs):
    world = int(np.prod(procs))
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _multi_gpu_worker import CASES, MODES
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
           str(world), "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_multi_gpu_worker.py"), out, ",".join(map(str, procs))]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    if p.returncode != 0:
        import glob
        errs = "".join(open(f).read()[-1500:] for f in sorted(glob.glob(out + ".err.*"))[:2])
        raise AssertionError(errs + p.stderr[-1500:])
    res = [json.load(open(f"{out}.{r}")) for r in range(world)]
    assert res[0]["pressure_cg@0"]["p2p"] == 1, "peer-memory path not active on an NVLink box"
    assert res[0]["pressure_cg@1"]["p2p"] == 0 and res[0]["pressure_cg@2"]["p2p"] == 1
    assert res[0]["pressure_cg@3"]["p2p"] == 1 and res[0]["pressure_cg@4"]["p2p"] == 1
    for name, (builder, solver, precond, mbs, tol) in (
            (f"{n}@{m}", c) for n, c in CASES.items() for m in MODES):
        systems = builder(procs)
        asms = [oracle.assemble(s) for s in systems]
        o = oracle.solve(asms, solver, precond, max_block_size=mbs, tolerance=tol, krylov_dim=30)
        iters = {r[name]["iters"] for r in res}
        assert len(iters) == 1, (name, iters)
        assert abs(iters.pop() - o.n_iterations) <= 2, name
        assert res[0][name]["global_n"] == sum(s.n for s in systems)
        x = gather_global(systems, [np.array(r[name]["x"]) for r in res])
        xo = gather_global(systems, o.x)
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-8, name
        assert res[0][name]["init"] == pytest.approx(o.init_residual, rel=1e-10)
        # distributed SpMV is bit-exact (same per-row order: local block, then halo entries)
        xg = np.random.default_rng(9).normal(size=sum(s.n for s in systems))
        ys = oracle.dist_spmv(asms, [xg[s.global_ids] for s in systems])
        for r in range(world):
            assert np.array_equal(np.array(res[r][name]["y"]), ys[r]), name
