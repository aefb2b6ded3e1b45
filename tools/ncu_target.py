"""Short PCG run for ncu (graphs off): python tools/ncu_target.py CELLS [opt=val ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402

cells = int(sys.argv[1])
s = bench.build_rank_system(cells, 1, 0)
ctx = Context()
ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
ctx.values_update(s.diag, s.upper)
ctx.vector_upload(L.OGL_VEC_B, s.source)
ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
ctx.set_option("use_graph", 0)
for a in sys.argv[2:]:
    k, v = a.split("=")
    ctx.set_option(k, int(v))
ctx.vector_fill(L.OGL_VEC_X, 0.0)
ctx.pcg_bench(12)
ctx.close()
