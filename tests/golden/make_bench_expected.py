"""Generates tests/golden/bench_expected.json: the oracle's iteration counts, initial residual and
norm factor for the workloads bench.py times (run here, once; takes a while at 64 M cells).

    python tests/golden/make_bench_expected.py [key ...]

Single-threaded oracle = the Ginkgo reference-executor summation order.  bench.py compares the GPU
run against these numbers in its `check` record (iterations within +-2), so the benchmark cannot
print a throughput for a wrong solve.  The GPU box has no /root/reference and little time: that is
why the counts are pinned here instead of being recomputed there."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from ogl_b200 import cases  # noqa: E402

OUT = os.environ.get("OGL_EXPECTED_OUT", os.path.join(ROOT, "tests", "golden", "bench_expected.json"))


def pressure(n, gpus):
    systems = [bench.build_rank_system(n, gpus, r) for r in range(gpus)]
    asms = [oracle.assemble(s) for s in systems]
    del systems
    return oracle.solve(asms, "GKOCG", "BJ", tolerance=bench.TOL, rel_tol=0.0, max_iter=bench.MAX_ITER, threads=1)


def pressure_ic(n):
    """The SPD twin of the pressure system (`scaling -1`, README.md:101) under GKOCG + IC."""
    a = oracle.assemble(bench.build_rank_system(n, 1, 0))
    a.vals, a.b = -a.vals, -a.b
    return oracle.solve([a], "GKOCG", "IC", tolerance=bench.TOL, rel_tol=0.0, max_iter=bench.MAX_ITER, threads=1)


def pressure_mg(n):
    a = oracle.assemble(bench.build_rank_system(n, 1, 0))
    a.vals, a.b = -a.vals, -a.b
    return oracle.solve([a], "GKOCG", "Multigrid", tolerance=bench.TOL, rel_tol=0.0, max_iter=bench.MAX_ITER, threads=1)


def momentum(n, precond="BJ"):
    s = cases.momentum_3d(n)[0]
    return oracle.solve([oracle.assemble(s)], "GKOBiCGStab", precond, tolerance=1e-5, rel_tol=0.0,
                        max_iter=2000, threads=1)


def channel(dims, procs):
    systems = cases.channel(dims, procs)
    return oracle.solve([oracle.assemble(s) for s in systems], "GKOGMRES", "BJ", tolerance=1e-6, rel_tol=0.0,
                        krylov_dim=100, threads=1)


JOBS = {
    "pressure_100_x1": lambda: pressure(100, 1),
    "pressure_200_x1": lambda: pressure(200, 1),
    "pressure_200_x2": lambda: pressure(200, 2),
    "pressure_200_x4": lambda: pressure(200, 4),
    "pressure_200_x8": lambda: pressure(200, 8),
    "pressure_100_x2": lambda: pressure(100, 2),
    "pressure_100_x4": lambda: pressure(100, 4),
    "pressure_100_x8": lambda: pressure(100, 8),
    "momentum_200_x1": lambda: momentum(200),
    "pressure_200_x1_ic": lambda: pressure_ic(200),
    "momentum_200_x1_ilu": lambda: momentum(200, "ILU"),
    "pressure_100_x1_mg": lambda: pressure_mg(100),
    "channel_128x64x64_x1": lambda: channel((128, 64, 64), (1, 1, 1)),
}

if __name__ == "__main__":
    keys = sys.argv[1:] or list(JOBS)
    for k in keys:
        t = time.time()
        r = JOBS[k]()
        rec = {"iterations": int(r.n_iterations), "criterion_calls": int(r.criterion_calls),
               "init_residual": float(r.init_residual), "final_residual": float(r.final_residual),
               "norm_factor": float(r.norm_factor), "oracle_seconds": round(time.time() - t, 1)}
        data = json.load(open(OUT)) if os.path.exists(OUT) else {}
        data[k] = rec
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
        print(k, rec, flush=True)
