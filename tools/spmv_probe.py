"""Run a few SpMVs of one kernel variant (for ncu):  python tools/spmv_probe.py <cells> <variant> [stages] [fused]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402

cells, variant = int(sys.argv[1]), int(sys.argv[2])
stages = int(sys.argv[3]) if len(sys.argv) > 3 else 3
fused = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
s = cases.pressure_3d(cells)[0]
ctx = Context()
ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
ctx.values_update(s.diag, s.upper)
ctx.set_option("spmv_variant", variant)
ctx.set_option("tma_stages", stages)
ms = ctx.spmv_bench(20, fused)
print(f"variant {variant} stages {stages} fused {fused}: {ms / 20 * 1e3:.2f} us")
