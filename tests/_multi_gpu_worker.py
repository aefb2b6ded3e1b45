"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.
Each rank solves its share through the plugin surface and dumps its results."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ogl_b200 import cases  # noqa: E402
from ogl_b200.host import ObjectRegistry  # noqa: E402
from ogl_b200.parallel import init_from_env  # noqa: E402
from ogl_b200.plugin import lduMatrix_solver_New  # noqa: E402

CASES = {
    "pressure_cg": (lambda p: cases.pressure_3d(16, p), "GKOCG", "BJ", 1, 1e-9),
    "pressure_cg_bj4": (lambda p: cases.pressure_3d(12, p), "GKOCG", "BJ", 4, 1e-9),
    "momentum_bicgstab": (lambda p: cases.momentum_3d(14, p), "GKOBiCGStab", "BJ", 1, 1e-10),
    # Schwarz-wrapped approximate inverses of the local blocks (no communication in the apply)
    "pressure_cg_isai": (lambda p: cases.pressure_3d(12, p, sign=-1.0), "GKOCG", "ISAI", 1, 1e-9),
    "momentum_bicgstab_gisai": (lambda p: cases.momentum_3d(10, p), "GKOBiCGStab", "GISAI", 1, 1e-10),
    # ... and incomplete factorisations of the local blocks (exact triangular sweeps)
    "pressure_cg_ic": (lambda p: cases.pressure_3d(12, p, sign=-1.0), "GKOCG", "IC", 1, 1e-9),
    "momentum_bicgstab_ilu": (lambda p: cases.momentum_3d(10, p), "GKOBiCGStab", "ILU", 1, 1e-10),
    "pressure_cg_multigrid": (lambda p: cases.pressure_3d(12, p, sign=-1.0), "GKOCG", "Multigrid", 1, 1e-9),
    "channel_gmres": (lambda p: cases.channel((16, 8, 8), p), "GKOGMRES", "BJ", 1, 1e-8),
    # all-Neumann + one reference cell: nearly singular, so solve tighter than the L2 bar
    # 2-D case: fold the z split into x ([2,2,2] -> [4,2,1] like test/integration.yaml:53-55)
    "cavity_cg_none": (lambda p: cases.cavity_2d((p[0] * p[2], p[1], 1)), "GKOCG", "none", 1, 1e-11),
}

MODES = (0, 1, 2, 3, 4, 5)
SHARED_MODES = (0, 2, 3, 4, 5)    # ranks sharing one device: no NCCL data path (mode 1)
BIG = 72                          # cells per direction per rank of the "beyond toy size" case


def big_dims(procs):
    return tuple(BIG * p for p in procs)


def big_rank_system(procs, rank):
    """>= 64^3 cells per rank: the automatic kernel selection of the benchmark sizes (pattern-coded
    ELL copy of the ghosted matrix, ghost-p CG, no persistent loop kernel)."""
    dims = big_dims(procs)
    return cases.build_rank_system(cases.PressureModel(dims, coef=1e-5), dims, procs, rank)


def main():
    out, procs = sys.argv[1], tuple(int(v) for v in sys.argv[2].split(","))
    shared = len(sys.argv) > 3 and sys.argv[3] == "shared"
    # shared: every rank on ONE device, gloo for the rendezvous and the window bootstrap
    ps = init_from_env("gloo", nccl=False) if shared else init_from_env("nccl")
    db = ObjectRegistry()
    results = {}
    modes = SHARED_MODES if shared else MODES
    # every case on the four data paths: 0 peer-memory windows, halo fused into the SpMV, CG
    # in ghost-p mode (default); 1 NCCL; 2 peer-memory windows with separate pack / non-local
    # kernels; 3 like 0 but CG with the flag handshake instead of ghost p; 4 / 5 the large-system
    # configuration forced onto these small cases: pattern-coded ELL copy of the ghosted matrix,
    # no persistent loop kernel, CG with (4) / without (5) the p-update fused into the SpMV (ghost
    # z pulled from the window inside the SpMV)
    for name, (builder, solver, precond, mbs, tol), mode in (
            (n, c, m) for n, c in CASES.items() for m in modes):
        s = builder(procs)[ps.rank]
        controls = {"solver": solver, "executor": "cuda", "tolerance": tol, "relTol": 0.0,
                    "adaptMinIter": False, "krylovDim": 30, "comm_mode": 1 if mode == 1 else 0,
                    "fused_halo": 0 if mode == 2 else 1, "ghost_p": 0 if mode == 3 else 1,
                    "preconditioner": {"preconditioner": precond, "maxBlockSize": mbs}}
        if mode >= 4:
            controls.update({"spmv_variant": 7, "ell_coded": 2, "fused_pcg": 0, "fuse_p": 1 if mode == 4 else 0,
                             "ell_tma": 1 if mode == 5 else 0})   # 5: the TMA-fed kernel over the ghosted matrix
        name = f"{name}@{mode}"
        sol = lduMatrix_solver_New(name, s, controls, db, ps)
        psi = s.psi.copy()
        perf = sol.solve(psi, s.source)
        x = np.random.default_rng(9).normal(size=sum(t.n for t in builder(procs)))[s.global_ids]
        y = sol.ctx.spmv(x)
        results[name] = {"iters": perf.n_iterations, "init": perf.initial_residual,
                         "final": perf.final_residual, "x": psi.tolist(), "y": y.tolist(),
                         "global_n": sol.ctx.partition_sizes()[1],
                         "p2p": sol.ctx.get_option("p2p_active")}
    # the benchmark-size configuration with everything on automatic
    s = big_rank_system(procs, ps.rank)
    controls = {"solver": "GKOCG", "executor": "cuda", "tolerance": 1e-7, "relTol": 0.0, "adaptMinIter": False,
                "preconditioner": "BJ"}
    sol = lduMatrix_solver_New("big", s, controls, db, ps)
    psi = s.psi.copy()
    perf = sol.solve(psi, s.source)
    results["big"] = {"iters": perf.n_iterations, "init": perf.initial_residual, "final": perf.final_residual,
                      "variant": sol.ctx.get_option("spmv_variant_in_use"),
                      "coded": sol.ctx.get_option("ell_coded_active"),
                      "fused_loop": sol.ctx.get_option("fused_pcg_active")}
    np.save(f"{out}.big.{ps.rank}.npy", psi)
    json.dump(results, open(f"{out}.{ps.rank}", "w"))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception:
        import traceback
        with open(f"{sys.argv[1]}.err.{os.environ.get('RANK', '0')}", "w") as f:
            traceback.print_exc(file=f)
        traceback.print_exc()
        raise
