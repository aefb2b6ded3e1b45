"""ELL (variant 7) vs pipelined CSR (variant 6): back-to-back SpMV and PCG iteration times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [100, 200]:
    s = cases.pressure_3d(n)[0]
    ctx = Context()
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
    ctx.values_update(s.diag, s.upper)
    ctx.vector_upload(L.OGL_VEC_B, s.source)
    ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
    b_spmv = 12 * ctx.nnz + 4 * (s.n + 1) + 16 * s.n
    for variant in (6, 7, 6, 7):
        ctx.set_option("spmv_variant", variant)
        ctx.spmv_bench(20, True)
        t0 = ctx.spmv_bench(200, False) / 200 * 1e3
        t1 = ctx.spmv_bench(200, True) / 200 * 1e3
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.pcg_bench(64)
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        us = ctx.pcg_bench(400) * 1e3 / 400
        print(json.dumps(dict(n=n, variant=variant, spmv_us=round(t0, 2), fused_us=round(t1, 2),
                              spmv_gbs=round(b_spmv / t0 / 1e3, 1), pcg_us=round(us, 2))), flush=True)
    ctx.close()
