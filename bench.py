#!/usr/bin/env python
"""Benchmark of the hot path: Jacobi-preconditioned CG on a synthetic 3-D
pressure system (BASELINE.json: "PCG iterations/sec & SpMV HBM GB/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells 100]

A *step* is one linear solve of the workload: GKOCG + BJ (maxBlockSize 1),
tolerance 1e-6, relTol 0, x0 = 0, on the N^3 lid-driven-cavity pressure matrix
(N = 100 per GPU: BASELINE configs[1]; `simple` box decomposition for --gpus > 1,
weak scaling).  Reported on one JSON line:

  value      PCG iterations/s with the system resident in HBM (coefficients,
             rhs and structure already on the device when the clock starts).
             For N GPUs: iterations x N blocks / s (whole-job aggregate of
             1M-cell block iterations); `global_iter_per_s` is the plain rate.
  e2e        the same metric through the plugin surface with HOST buffers:
             every step uploads the LDU coefficients + rhs + initial guess from
             pinned memory, solves, and downloads the solution.
  roofline   FP64 CSR SpMV fused with <p,q> (the dominant kernel): algorithmic
             bytes 12 nnz + 4 (n+1) + 16 n per launch / CUDA-event duration,
             against the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline  the oracle's PCG (port of the reference's Ginkgo reference
             executor path) on a bounded sample of the same system.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-6
MAX_ITER = 1000


def procs_for(n_gpus: int):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n_gpus]


def alg_bytes_spmv(n, nnz, n_halo=0):
    # CSR (8 B value + 4 B column) + row pointers + x + y
    return 12 * nnz + 4 * (n + 1) + 16 * n + (12 * n_halo + 8 * n_halo + 12 * n_halo)


def alg_bytes_pcg(n, nnz, n_halo=0):
    # fused-minimum traffic of one Jacobi-PCG iteration
    return 12 * nnz + 4 * (n + 1) + 96 * n + (12 * n_halo + 8 * n_halo + 12 * n_halo)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(cells: int):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return rec.get(str(cells), {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", delete=False, suffix=".csv")
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nme, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def build_rank_system(n: int, n_gpus: int, rank: int):
    from ogl_b200 import cases
    px, py, pz = procs_for(n_gpus)
    dims = (n * px, n * py, n * pz)
    m = cases.PressureModel(dims, coef=1e-5)
    return cases.build_rank_system(m, dims, (px, py, pz), rank)


# ----------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# ----------------------------------------------------------------------------

def cpu_pcg_sample(system, threads: int, iters: int):
    """`iters` PCG iterations of the oracle on `system`; returns (it/s, seconds)."""
    import oracle
    a = oracle.assemble(system)
    r = oracle.solve([a], "GKOCG", "BJ", tolerance=0.0, rel_tol=0.0, max_iter=iters,
                     threads=threads)
    done = max(r.criterion_calls - 1, 1)
    return done / r.seconds, r.seconds, done


def cpu_foam_sample(system, iters: int):
    """`iters` iterations of the OpenFOAM-native-equivalent PCG (face-based Amul + diagonal
    preconditioner, oracle/foam_pcg.cpp), single thread; returns (it/s, seconds, iterations)."""
    import oracle
    r = oracle.foam_pcg(system, tolerance=0.0, rel_tol=0.0, max_iter=iters)
    done = max(r.n_iterations, 1)
    return done / r.seconds, r.seconds, done


def workload_text(n: int) -> str:
    return (f"{n}^3 cells per GPU, 3-D lid-driven-cavity pressure system, GKOCG+BJ(maxBlockSize 1) FP64, "
            "tolerance 1e-6, relTol 0, x0=0 (BASELINE configs[1])")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    s = build_rank_system(args.n, 1, 0)
    # bounded sample per step so that steps+warmup end within a few minutes
    sample_iters = 60
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        rate, sec, done = cpu_pcg_sample(s, cores, sample_iters)
        if i >= args.warmup:
            rates.append(rate)
            secs.append(sec)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "PCG iterations/sec", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.n),
                   "rows_per_gpu": s.n, "nnz_per_gpu": s.n + 2 * s.n_faces,
                   "value_definition": "iterations x n_gpus / s (1M-cell block iterations, whole job): the "
                                       "host cores work through the blocks one after the other, so the "
                                       "figure does not depend on n_gpus"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_iters} PCG iterations per step of the oracle port "
                                   f"(OpenMP, {cores} threads) on the full {args.n}^3 system; the "
                                   "reference itself (OGL+Ginkgo+OpenFOAM) cannot be built here"},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist

    from ogl_b200 import _lib as L
    from ogl_b200 import host
    from ogl_b200.backend import Context
    from ogl_b200.parallel import init_from_env

    ps = init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    n_gpus = ps.n_ranks
    if n_gpus != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={n_gpus} (launch with torchrun)")
    torch.cuda.set_device(ps.local_rank)
    dev = torch.device("cuda", ps.local_rank)
    stream = torch.cuda.current_stream()

    s = build_rank_system(args.n, n_gpus, ps.rank)
    n, nf = s.n, s.n_faces
    ctx = Context(device_id=ps.local_rank, rank=ps.rank, n_ranks=n_gpus, nccl_id=ps.nccl_id,
                  stream=stream.cuda_stream)
    # ---- setup (untimed): structure goes to the device once and stays there
    ir, ic = host.collect_local_interface_indices(s)
    ctx.pattern_from_ldu(n, s.lower_addr, s.upper_addr, True, ir, ic)
    ctx.partition_create(n, *host.create_communication_pattern(s))
    ctx.nonlocal_pattern(host.collect_cells_on_non_local_interface(s))
    nnz, n_halo = ctx.nnz, ctx.n_halo
    p2p_active = bool(ctx.get_option("p2p_active"))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
    h_diag, h_upper, h_b, h_x0 = pin(s.diag), pin(s.upper), pin(s.source), pin(s.psi)
    h_if = pin(host.collect_interface_coeffs(s, True))
    h_nl = pin(host.collect_interface_coeffs(s, False))
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def upload_values():
        ctx.values_update(h_diag, h_upper, None, h_if if h_if.numel() else None,
                          h_nl if h_nl.numel() else None, 1.0)

    def barrier():
        torch.cuda.synchronize()
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve():
        ctx.precond_setup(L.OGL_PRECOND_BJ, 1)
        return ctx.solve(L.OGL_SOLVER_CG, tolerance=TOL, rel_tol=0.0, max_iter=MAX_ITER)

    def step_resident():
        ctx.vector_fill(L.OGL_VEC_X, 0.0)
        return solve()

    def step_e2e():
        upload_values()
        ctx.vector_upload(L.OGL_VEC_B, h_b)
        ctx.vector_upload(L.OGL_VEC_X, h_x0)
        r = solve()
        ctx.vector_download(L.OGL_VEC_X, h_out)
        return r

    def timed(step_fn, steps, warmup):
        """K steps, each bracketed by CUDA events on the launching stream, L2
        flushed between steps; max over ranks of the summed step times."""
        for _ in range(warmup):
            flush.zero_()
            step_fn()
        barrier()
        launches0 = ctx.get_option("launches")
        ms, iters = 0.0, 0
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = step_fn()
            e1.record(stream)
            e1.synchronize()
            ms += e0.elapsed_time(e1)
            iters += r.n_iterations
        barrier()
        launches = ctx.get_option("launches") - launches0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if n_gpus > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), iters, launches

    upload_values()
    ctx.vector_upload(L.OGL_VEC_B, h_b)
    sampler = ClockSampler(ps.local_rank)
    if ps.rank == 0:
        sampler.start()
    ms_res, iters_res, launches = timed(step_resident, args.steps, args.warmup)
    ms_e2e, iters_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup - 2))
    clocks = sampler.stop() if ps.rank == 0 else {}

    # ---- dominant kernel: SpMV fused with <p,q>.  Measured live in the REAL CG loop:
    # one more (untimed) solve with the chunk graph off and CUDA events recorded on the
    # launching stream around every 4th SpMV launch; plus the back-to-back figures
    ctx.set_option("use_graph", 0)
    ctx.set_option("profile_stride", 4)
    flush.zero_()
    r_prof = step_resident()
    ctx.set_option("profile_stride", 0)
    ctx.set_option("use_graph", 1)
    reps = 200
    spmv_b2b_ms = ctx.spmv_bench(reps, fused_dot=True) / reps
    spmv_plain_ms = ctx.spmv_bench(reps, fused_dot=False) / reps
    spmv_ms = r_prof.spmv_us_avg * 1e-3 if r_prof.spmv_samples > 0 else spmv_b2b_ms
    t = torch.tensor([spmv_ms, spmv_plain_ms, spmv_b2b_ms], dtype=torch.float64, device=dev)
    if n_gpus > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    spmv_ms, spmv_plain_ms, spmv_b2b_ms = (float(v) for v in t.tolist())

    if ps.rank != 0:
        ctx.close()
        return 0

    variant = ctx.get_option("spmv_variant_in_use")
    peak, peak_src = measured_peak()
    b_spmv = alg_bytes_spmv(n, nnz, n_halo)
    b_pcg = alg_bytes_pcg(n, nnz, n_halo)
    achieved = b_spmv / (spmv_ms * 1e-3) / 1e9
    it_per_s = iters_res / (ms_res * 1e-3)
    pcg_gbs = b_pcg * it_per_s / 1e9
    e2e_it_per_s = iters_e2e / (ms_e2e * 1e-3)
    h2d = 8 * nf + 8 * n + 8 * n + 8 * n + 8 * (h_if.numel() + h_nl.numel())
    d2h = 8 * n

    # ---- CPU baseline: the oracle port, single thread (reference-executor order);
    # on rank 0 at N = 1 only (a rank's block of a decomposed case is not a closed system)
    cpu_baseline = None
    if n_gpus == 1:
        cpu_iters = 40 if args.n >= 100 else 200
        cpu_rate, cpu_sec, cpu_done = cpu_pcg_sample(s, 1, cpu_iters)
        foam_rate, foam_sec, foam_done = cpu_foam_sample(s, cpu_iters)
        cpu_baseline = {"value": cpu_rate, "unit": "iter/s", "cores": 1, "kind": "port",
                        "sample": f"{cpu_done} PCG iterations of the oracle (single thread, Ginkgo "
                                  f"reference-executor order) on the full {args.n}^3 system, "
                                  f"{cpu_sec:.1f} s",
                        # what OpenFOAM itself would run without OGL (baseline iii)
                        "openfoam_native_equivalent": {
                            "value": foam_rate, "unit": "iter/s", "cores": 1, "kind": "port",
                            "sample": f"{foam_done} iterations of face-based lduMatrix::Amul + diagonal PCG "
                                      f"(oracle/foam_pcg.cpp) on the same system, {foam_sec:.1f} s"}}

    line = {
        "metric": "PCG iterations/sec", "value": it_per_s * n_gpus, "unit": "iter/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_text(args.n),
            "rows_per_gpu": n, "nnz_per_gpu": nnz, "halo_per_gpu": n_halo,
            "decomposition": list(procs_for(n_gpus)),
            "comm": ("none" if n_gpus == 1 else
                     ("peer-memory windows over NVLink (stamped P2P stores of the boundary z + in-kernel "
                      "all-reduce, no halo handshake in the CG loop)"
                      if p2p_active else "NCCL send/recv + allreduce")),
            "iterations_per_solve": iters_res / args.steps,
            "l2": "working set ~%.0f MB vs 126 MB L2; 512 MB written between steps to flush L2; "
                  "inside a solve the iterations reuse whatever L2 keeps (that is the workload)"
                  % ((12 * nnz + 4 * n + 5 * 8 * n) / 1e6),
            "value_definition": "iterations x n_gpus / s (1M-cell block iterations, whole job)",
        },
        "global_iter_per_s": it_per_s,
        "e2e": {"value": e2e_it_per_s * n_gpus, "unit": "iter/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": ("k_spmv_ell<false,1> (FP64 ELL SpMV + fused <p,q>)" if variant == 7 else
                                       "k_spmv_pipe<false,1,false> (FP64 CSR SpMV + fused <p,q>%s)" % (
                "; ghosted CSR, all-reduce of <p,q> inside the launch" if n_gpus > 1 else "")),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # the committed ncu capture is of the CSR kernel
            "traffic": ncu_traffic(args.n) if n_gpus == 1 and variant != 7 else None, "peak_source": peak_src,
            "alg_bytes_per_launch": b_spmv, "us_per_launch": spmv_ms * 1e3,
            "us_per_launch_how": ("CUDA events around every 4th SpMV launch inside an extra "
                                  "solve of the same system (%d samples)" % r_prof.spmv_samples),
            "us_per_launch_back_to_back": spmv_b2b_ms * 1e3,
            "us_per_launch_unfused": spmv_plain_ms * 1e3,
            "frac_of_nominal_8TBs": achieved / 8000.0,
            "pcg_iteration": {"alg_bytes": b_pcg, "achieved": pcg_gbs, "frac": pcg_gbs / peak,
                              "us_per_iteration": 1e6 / it_per_s},
            "note": "1M rows: matrix+vectors ~ L2 size, so achieved GB/s is not a clean HBM "
                    "figure (see extra / profiles for 200^3)",
        },
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ogl_b200", choices=["ogl_b200", "reference"])
    ap.add_argument("--cells", dest="n", type=int, default=100, help="cells per direction per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.gpus not in (1, 2, 4, 8):
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
