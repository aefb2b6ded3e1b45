"""A/B of the L2-residency policy of the matrix stream (run on the GPU box):
    python tools/l2_probe.py [cells ...]
For every `l2_keep_mb` setting: back-to-back SpMV and whole PCG iterations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100, 200]
    for n in sizes:
        s = cases.pressure_3d(n)[0]
        ctx = Context()
        ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
        ctx.values_update(s.diag, s.upper)
        ctx.vector_upload(L.OGL_VEC_B, s.source)
        ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
        nnz = ctx.nnz
        b_spmv = 12 * nnz + 4 * (s.n + 1) + 16 * s.n
        b_pcg = 12 * nnz + 4 * (s.n + 1) + 96 * s.n
        for keep in (0, -1, 32, 48, 64, 76, 100, 112, 0, -1):
            ctx.set_option("l2_keep_mb", keep)
            level = ctx.get_option("l2_keep_level")
            reps = 200
            ctx.spmv_bench(20, True)
            t0 = ctx.spmv_bench(reps, False) / reps * 1e3
            t1 = ctx.spmv_bench(reps, True) / reps * 1e3
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            ctx.pcg_bench(50)
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            iters = 400
            us = ctx.pcg_bench(iters) * 1e3 / iters
            print(json.dumps(dict(n=n, l2_keep_mb=keep, level=level, spmv_us=round(t0, 2), fused_us=round(t1, 2),
                                  spmv_gbs=round(b_spmv / t0 / 1e3, 1), pcg_us=round(us, 2),
                                  pcg_gbs=round(b_pcg / us / 1e3, 1))), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
