"""ILU / IC sweeps, IRILU sweeps and one multigrid V cycle for ncu: python tools/ncu_target3.py [cells]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s = cases.pressure_3d(cells)[0]
ctx = Context()
ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
ctx.values_update(s.diag, s.upper, scaling=-1.0)
r = np.random.default_rng(1).standard_normal(s.n)
for kind in (L.OGL_PRECOND_IC, L.OGL_PRECOND_IRILU, L.OGL_PRECOND_MULTIGRID):
    ctx.precond_setup(kind, 1)
    ctx.precond_apply(r)
    ctx.precond_apply(r)
ctx.close()
