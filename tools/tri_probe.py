"""ILU / IC / IRILU at the benchmark sizes: analysis + factorisation time, time per iteration of
the sweep variants and their knobs next to scalar Jacobi (one GPU).
Usage: python tools/tri_probe.py [cells ...]"""
import json
import os
import sys
import time

sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bench.py") else ".")
import bench
from ogl_b200 import _lib as L
from ogl_b200.backend import Context

CONFIGS = [("BJ", L.OGL_PRECOND_BJ, 1, 100, 0), ("IC", L.OGL_PRECOND_IC, 0, 100, 0)]
CONFIGS += [("IC", L.OGL_PRECOND_IC, 1, sl, ct) for sl, ct in ((0, 0), (100, 0), (300, 0), (1000, 0), (0, 1), (100, 1), (300, 1))]
CONFIGS += [("ILU", L.OGL_PRECOND_ILU, 1, 0, 0), ("IRILU", L.OGL_PRECOND_IRILU, 1, 0, 0),
            ("Multigrid", L.OGL_PRECOND_MULTIGRID, 1, 0, 0)]
if os.environ.get("TRI_PROBE_SHORT"):
    CONFIGS = [c for c in CONFIGS if c[0] in ("BJ", "Multigrid") or (c[0] == "IC" and c[2:] == (1, 0, 0))]

for cells in [int(v) for v in sys.argv[1:]] or [100, 200]:
    s = bench.build_rank_system(cells, 1, 0)
    ctx = Context()
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
    ctx.values_update(s.diag, s.upper, scaling=-1.0)     # SPD twin of the pressure matrix (README.md:101)
    ctx.vector_upload(L.OGL_VEC_B, s.source, -1.0)
    for name, kind, variant, sleep_ns, ctas in CONFIGS:
        ctx.set_option("tri_variant", variant)
        ctx.set_option("tri_sleep_ns", sleep_ns)
        ctx.set_option("tri_ctas", ctas)
        best = None
        setup_ms = []
        for _ in range(2):
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            ctx.synchronize()
            t0 = time.perf_counter()
            ctx.precond_setup(kind, 1)
            ctx.synchronize()
            setup_ms.append((time.perf_counter() - t0) * 1e3)
            solver = L.OGL_SOLVER_CG if name != "IRILU" else L.OGL_SOLVER_BICGSTAB
            r = ctx.solve(solver, tolerance=1e-6, max_iter=2000)
            if best is None or r.solve_us < best.solve_us:
                best = r
        print(json.dumps({"cells": cells, "precond": name, "tri_variant": variant, "sleep_ns": sleep_ns, "ctas": ctas,
                          "iters": best.n_iterations, "us_per_iter": round(best.solve_us / max(best.n_iterations, 1), 1),
                          "solve_ms": round(best.solve_us / 1e3, 2), "final": best.final_residual,
                          "setup_ms_first": round(setup_ms[0], 2), "setup_ms": round(setup_ms[1], 2),
                          "levels": ctx.get_option("tri_levels_lower"), "launches": best.kernel_launches,
                          "mg_levels": [(l["n"], l["nnz"]) for l in ctx.mg_levels()] if name == "Multigrid" else None}),
              flush=True)
    ctx.close()
