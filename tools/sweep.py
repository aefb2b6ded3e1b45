"""Timing sweep over the library's tuning options (run on the GPU box):
    python tools/sweep.py [cells ...]
SpMV (plain / fused) and whole PCG iterations via the C ABI bench entry points."""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import cases, host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100, 200]
    out = []
    for n in sizes:
        s = cases.pressure_3d(n)[0]
        ctx = Context()
        ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
        ctx.values_update(s.diag, s.upper)
        ctx.vector_upload(L.OGL_VEC_B, s.source)
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
        nnz = ctx.nnz
        print(json.dumps(dict(n=n, hbm_copy_gbs=round(ctx.membench(0), 1),
                              hbm_read_gbs=round(ctx.membench(1), 1),
                              hbm_read_8p4_gbs=round(ctx.membench(2), 1))), flush=True)
        b_spmv = 12 * nnz + 4 * (s.n + 1) + 16 * s.n
        b_pcg = 12 * nnz + 4 * (s.n + 1) + 96 * s.n
        for variant, ctas, stages, blocked in itertools.product((1, 6, 4), (0, 148 * 4), (2,), (0,)):
            ctx.set_option("tile_blocked", blocked)
            if variant == 4 and ctas:
                continue
            ctx.set_option("tma_stages", stages)
            try:
                ctx.set_option("spmv_variant", variant)
                ctx.set_option("stream_ctas", ctas)
            except Exception as e:
                print("skip", variant, ctas, e)
                continue
            reps = 100
            t0 = ctx.spmv_bench(reps, False) / reps * 1e3
            t1 = ctx.spmv_bench(reps, True) / reps * 1e3
            row = dict(n=n, variant=variant, blocked=blocked, stages=stages, spmv_us=round(t0, 2), fused_us=round(t1, 2),
                       spmv_gbs=round(b_spmv / t0 / 1e3, 1), fused_gbs=round(b_spmv / t1 / 1e3, 1))
            print(json.dumps(row), flush=True)
            out.append(row)
        ctx.set_option("spmv_variant", 0)
        ctx.set_option("stream_ctas", 0)
        for blocks, graph, chunk in itertools.product((148 * 4,), (1,), (16,)):
            ctx.set_option("blas1_blocks", blocks)
            ctx.set_option("use_graph", graph)
            ctx.set_option("chunk_iters", chunk)
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            iters = 400
            ctx.pcg_bench(50)
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            ms = ctx.pcg_bench(iters)
            us = ms * 1e3 / iters
            row = dict(n=n, blas1_blocks=blocks, graph=graph, chunk=chunk, pcg_us=round(us, 2),
                       pcg_gbs=round(b_pcg / us / 1e3, 1))
            print(json.dumps(row), flush=True)
            out.append(row)
        ctx.close()
    json.dump(out, open(os.path.join("gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
