"""Multi-GPU micro-benchmarks (launch with torchrun, one rank per GPU):
peer-memory all-reduce / halo exchange latency, SpMV and PCG iteration times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ogl_b200 import _lib as L  # noqa: E402
from ogl_b200 import host  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402
from ogl_b200.parallel import init_from_env  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    ps = init_from_env("nccl")
    s = bench.build_rank_system(cells, ps.n_ranks, ps.rank)
    out = {}
    for mode in (0, 3, 2, 1):
        ctx = Context(device_id=ps.local_rank, rank=ps.rank, n_ranks=ps.n_ranks, nccl_id=ps.nccl_id)
        ctx.set_option("comm_mode", 1 if mode == 1 else 0)
        ctx.set_option("fused_halo", 0 if mode == 2 else 1)
        ctx.set_option("ghost_p", 0 if mode == 3 else 1)
        ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
        ctx.partition_create(s.n, *host.create_communication_pattern(s))
        ctx.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
        ctx.values_update(s.diag, s.upper, None, None, host.collect_interface_coeffs(s, False))
        ctx.vector_upload(L.OGL_VEC_B, s.source)
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
        tag = ({0: "p2p_ghost_p", 3: "p2p_fused"}.get(mode, "p2p")) if ctx.get_option("p2p_active") else "nccl"
        r = {}
        if tag == "p2p":
            r["ar_per_launch_us"] = ctx.commbench(0, 300)
            r["ar_device_us"] = ctx.commbench(1, 2000)
            r["halo_exchange_us"] = ctx.commbench(2, 300)
        r["spmv_us"] = ctx.spmv_bench(200, False) / 200 * 1e3
        r["spmv_fused_us"] = ctx.spmv_bench(200, True) / 200 * 1e3
        ctx.pcg_bench(50)
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        r["pcg_iter_us"] = ctx.pcg_bench(400) / 400 * 1e3
        out[tag] = {k: round(v, 2) for k, v in r.items()}
        ctx.close()
    if ps.rank == 0:
        print(json.dumps({"cells": cells, "n_gpus": ps.n_ranks, **out}), flush=True)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
