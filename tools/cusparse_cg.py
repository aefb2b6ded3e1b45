"""Library comparator: the same Jacobi-PCG composed from cuSPARSE's CSR SpMV
(torch.sparse_csr @ dense -> cusparseSpMV, FP64) and library BLAS-1 (torch.dot / elementwise
kernels), with the same L1 / normFactor stopping rule evaluated the way the reference does it --
one device->host scalar per iteration.  NOT part of the product and not used by bench.py's
default run; it answers "what does the stock-library composition of this iteration cost on the
same box?".

    python tools/cusparse_cg.py [cells] [--device cpu]      # prints one JSON line
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogl_b200 import cases  # noqa: E402

SMALL = 1e-15


def csr_of(s, device):
    """Row-major CSR of one rank's lduMatrix (no interfaces), as torch.sparse_csr (FP64)."""
    n = s.n
    lower = s.upper if s.lower is None else s.lower
    rows = np.concatenate([np.arange(n), s.lower_addr, s.upper_addr]).astype(np.int64)
    cols = np.concatenate([np.arange(n), s.upper_addr, s.lower_addr]).astype(np.int64)
    vals = np.concatenate([s.diag, s.upper, lower])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    crow = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=crow[1:])
    return torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(cols), torch.from_numpy(vals),
                                   size=(n, n), dtype=torch.float64, device=device)


def pcg(A, b, x, inv_diag, tolerance=1e-6, max_iter=1000):
    """Jacobi-PCG in Ginkgo's cg.cpp order under OGL's criterion (StoppingCriterion.C:11-151):
    returns (criterion calls, initial residual, final residual, seconds in the loop)."""
    n = b.numel()
    ones = torch.ones_like(b)
    w = A @ (ones * x.mean())
    r = b - A @ x
    norm_factor = float((torch.abs(b - w) + torch.abs((b - w) - r)).sum()) + SMALL
    z = r * inv_diag
    rho = torch.dot(r, z)
    p = torch.zeros_like(b)
    prev_rho = None
    init = final = float(torch.abs(r).sum()) / norm_factor    # criterion call 0
    calls = 1
    if final < tolerance:
        return calls, init, final, 0.0
    if A.is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    while True:
        p = z.clone() if prev_rho is None else z + (rho / prev_rho) * p
        q = A @ p
        beta = torch.dot(p, q)
        alpha = rho / beta
        x += alpha * p
        r -= alpha * q
        z = r * inv_diag
        prev_rho = rho
        rho = torch.dot(r, z)
        final = float(torch.abs(r).sum()) / norm_factor      # D2H + sync every iteration
        calls += 1
        if final < tolerance or calls > max_iter:
            break
    if A.is_cuda:
        torch.cuda.synchronize()
    return calls, init, final, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cells", type=int, nargs="?", default=100)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    s = cases.pressure_3d(args.cells)[0]
    dev = torch.device(args.device)
    A = csr_of(s, dev)
    b = torch.from_numpy(s.source).to(dev)
    inv_diag = 1.0 / torch.from_numpy(s.diag).to(dev)
    best = None
    for _ in range(args.repeats):
        x = torch.zeros_like(b)
        calls, init, final, sec = pcg(A, b, x, inv_diag)
        if best is None or sec < best[3]:
            best = (calls, init, final, sec)
    calls, init, final, sec = best
    nnz = s.n + 2 * s.n_faces
    print(json.dumps({"comparator": "torch.sparse_csr (cusparseSpMV FP64) + library BLAS-1, host-checked criterion",
                      "cells": args.cells, "device": str(dev), "criterion_calls": calls,
                      "init_residual": init, "final_residual": final,
                      "us_per_iteration": 1e6 * sec / max(calls - 1, 1),
                      "iter_per_s": max(calls - 1, 1) / sec if sec > 0 else None,
                      "alg_gbs": (12 * nnz + 4 * (s.n + 1) + 96 * s.n) * max(calls - 1, 1) / sec / 1e9 if sec > 0 else None}))


if __name__ == "__main__":
    main()
