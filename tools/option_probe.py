"""A/B of one or more library options on the resident PCG loop (run on the GPU box; one
rank or torchrun):  python tools/option_probe.py CELLS opt=v1,v2[,..] [opt2=...]
Prints us per PCG iteration (ogl_pcg_bench, 400 iterations) for every combination."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def main():
    cells = int(sys.argv[1])
    opts = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[2:]]
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        from ogl_b200.parallel import init_from_env
        ps = init_from_env("nccl")
        s = bench.build_rank_system(cells, ps.n_ranks, ps.rank)
        ctx = Context(device_id=ps.local_rank, rank=ps.rank, n_ranks=ps.n_ranks, nccl_id=ps.nccl_id)
        rank, n_ranks = ps.rank, ps.n_ranks
    else:
        s = bench.build_rank_system(cells, 1, 0)
        ctx = Context()
        rank, n_ranks = 0, 1
    ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
    if multi:
        ctx.partition_create(s.n, *host.create_communication_pattern(s))
        ctx.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
        ctx.values_update(s.diag, s.upper, None, None, host.collect_interface_coeffs(s, False))
    else:
        ctx.values_update(s.diag, s.upper)
    ctx.vector_upload(L.OGL_VEC_B, s.source)
    ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
    b_pcg = 12 * ctx.nnz + 4 * (s.n + 1) + 96 * s.n
    for combo in itertools.product(*[v for _, v in opts]):
        for (k, _), v in zip(opts, combo):
            ctx.set_option(k, v)
        ctx.spmv_bench(20, True)
        spmv_us = min(ctx.spmv_bench(200, False) for _ in range(2)) / 200 * 1e3
        fused_us = min(ctx.spmv_bench(200, True) for _ in range(2)) / 200 * 1e3
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.pcg_bench(64)
        best = []
        for _ in range(3):
            ctx.vector_fill(L.OGL_VEC_X, 0.0)
            iters = 400
            best.append(ctx.pcg_bench(iters) * 1e3 / iters)
        if rank == 0:
            us = min(best)
            print(json.dumps({"cells": cells, "n_gpus": n_ranks, **{k: v for (k, _), v in zip(opts, combo)},
                              "fused_active": ctx.get_option("fused_pcg_active"),
                              "variant": ctx.get_option("spmv_variant_in_use"),
                              "coded": ctx.get_option("ell_coded_active"),
                              "patterns": [ctx.get_option("ell_patterns"), ctx.get_option("gell_patterns")],
                              "escapes": [ctx.get_option("ell_escape_rows"), ctx.get_option("gell_escape_rows")],
                              "spmv_us": round(spmv_us, 2), "spmv_dot_us": round(fused_us, 2),
                              "pcg_us": round(us, 2), "pcg_us_all": [round(b, 2) for b in best],
                              "pcg_gbs": round(b_pcg / us / 1e3, 1)}), flush=True)
    ctx.close()
    if multi:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
