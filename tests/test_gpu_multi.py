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


def run_workers(tmp_path, procs, shared):
    world = int(np.prod(procs))
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
           str(world), "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_multi_gpu_worker.py"), out, ",".join(map(str, procs))]
    env = dict(os.environ)
    if shared:
        cmd.append("shared")
        env["OGL_B200_DEVICE"] = "0"
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    if p.returncode != 0:
        import glob
        errs = "".join(open(f).read()[-1500:] for f in sorted(glob.glob(out + ".err.*"))[:2])
        raise AssertionError(errs + p.stderr[-1500:])
    return out, [json.load(open(f"{out}.{r}")) for r in range(world)]


# `shared`: all ranks of the decomposition on device 0 (one process each, CUDA-IPC windows inside
# the device, gloo for the bootstrap) -- the decomposed path runs on a single-GPU box too.
def _variants():
    """Ranks sharing device 0 always; one rank per GPU for every decomposition the box has GPUs for
    (collected, not skipped: a 1-GPU box runs the whole multi-rank code through the shared variants)."""
    out = [((2, 1, 1), True), ((2, 2, 1), True)]
    try:
        have = n_gpus()
    except Exception:
        have = 0
    out += [(p, False) for p in ((2, 1, 1), (2, 2, 1), (2, 2, 2)) if int(np.prod(p)) <= have]
    return out


@pytest.mark.parametrize("procs,shared", _variants())
def test_decomposed_solves_match_oracle(oracle, tmp_path, procs, shared):
    world = int(np.prod(procs))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _multi_gpu_worker import CASES, MODES, SHARED_MODES, big_dims
    from ogl_b200 import cases
    out, res = run_workers(tmp_path, procs, shared)
    modes = SHARED_MODES if shared else MODES
    assert res[0]["pressure_cg@0"]["p2p"] == 1, "peer-memory path not active"
    if not shared:
        assert res[0]["pressure_cg@1"]["p2p"] == 0
    assert res[0]["pressure_cg@2"]["p2p"] == 1
    assert res[0]["pressure_cg@3"]["p2p"] == 1 and res[0]["pressure_cg@4"]["p2p"] == 1
    # the benchmark-size case: automatic selection = coded ELL copy of the ghosted matrix + ghost-p CG
    dims = big_dims(procs)
    systems = cases.build_case(cases.PressureModel(dims, coef=1e-5), dims, procs)
    o = oracle.solve([oracle.assemble(s) for s in systems], "GKOCG", "BJ", tolerance=1e-7,
                     threads=os.cpu_count() or 1)
    assert {r["big"]["variant"] for r in res} == {7} and all(r["big"]["coded"] & 2 for r in res)
    assert all(r["big"]["fused_loop"] == 0 for r in res)
    assert {abs(r["big"]["iters"] - o.n_iterations) <= 2 for r in res} == {True}
    x = gather_global(systems, [np.load(f"{out}.big.{r}.npy") for r in range(world)])
    xo = gather_global(systems, o.x)
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) <= 1e-8
    del systems, o, x, xo
    for name, (builder, solver, precond, mbs, tol) in (
            (f"{n}@{m}", c) for n, c in CASES.items() for m in modes):
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
