"""Device-side timeline of the PCG iteration (run on the GPU box; one rank or torchrun):
    python tools/trace_iter.py [cells] [option=value ...]
The kernels log %globaltimer at launch start / last CTA arrived / local sums done /
all-reduced / epilogue done (option `trace`); this prints the mean time between
consecutive events of the steady-state iterations."""
import collections
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402

NAMES = {10: "p.start", 20: "spmv.start", 21: "spmv.last_cta", 22: "spmv.sums", 23: "spmv.allreduced",
         24: "spmv.epi", 30: "xr.start", 31: "xr.last_cta", 32: "xr.sums", 33: "xr.allreduced", 34: "xr.epi"}


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    opts = dict(a.split("=") for a in sys.argv[2:])
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        from ogl_b200.parallel import init_from_env
        ps = init_from_env("nccl")
        s = bench.build_rank_system(cells, ps.n_ranks, ps.rank)
        ctx = Context(device_id=ps.local_rank, rank=ps.rank, n_ranks=ps.n_ranks, nccl_id=ps.nccl_id)
        rank = ps.rank
    else:
        s = bench.build_rank_system(cells, 1, 0)
        ctx = Context()
        rank = 0
    for k, v in opts.items():
        ctx.set_option(k, int(v))
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
    if multi:
        ctx.partition_create(s.n, *host.create_communication_pattern(s))
        ctx.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
        ctx.values_update(s.diag, s.upper, None, None, host.collect_interface_coeffs(s, False))
    else:
        ctx.values_update(s.diag, s.upper)
    ctx.vector_upload(L.OGL_VEC_B, s.source)
    ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
    ctx.vector_fill(L.OGL_VEC_X, 0.0)
    ctx.pcg_bench(64)
    ctx.set_option("trace", 1)
    ctx.vector_fill(L.OGL_VEC_X, 0.0)
    iters = 160
    us = ctx.pcg_bench(iters) * 1e3 / iters
    tags, times = ctx.trace_download()
    order = np.argsort(times, kind="stable")
    tags, times = tags[order], times[order]
    # steady state: drop the first and last 2 chunks
    first_tag = 10 if np.any(tags == 10) else 20      # fuse_p: no p-update kernel, the SpMV opens the iteration
    starts = np.flatnonzero(tags == first_tag)
    lo, hi = starts[32], starts[-32]
    gaps = collections.defaultdict(list)
    for i in range(lo, hi):
        gaps[(int(tags[i]), int(tags[i + 1]))].append(times[i + 1] - times[i])
    rows = {f"{NAMES.get(a, a)}->{NAMES.get(b, b)}": (round(float(np.mean(v)) / 1e3, 2), len(v))
            for (a, b), v in sorted(gaps.items())}
    total = sum(m * c for m, c in rows.values()) / max(1, len(starts[32:-32]))
    print(json.dumps({"rank": rank, "cells": cells, "opts": opts, "pcg_iter_us_traced": round(us, 2),
                      "sum_of_gaps_us": round(total, 2), "gaps_us(mean,count)": rows}), flush=True)
    ctx.close()
    if multi:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
