"""Generates tests/golden/precond_oracle_vectors.npz: REGRESSION fixtures of the oracle's own
preconditioner restatements (oracle/trifactor.hpp, oracle/multigrid.hpp) on tiny systems.

These vectors pin the oracle against accidental change -- they do NOT pin it against the reference:
the algorithms live in Ginkgo, which is absent here (parity unpinned, see oracle/krylov.cpp).

    python tests/golden/make_precond_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ogl_b200 import cases  # noqa: E402


def csr(s):
    a = oracle.assemble(s)
    rp = np.zeros(s.n + 1, np.int32)
    np.cumsum(np.bincount(a.rows, minlength=s.n), out=rp[1:])
    return a, rp


def build():
    out = {}
    s = cases.momentum_3d(5)[0]
    a, rp = csr(s)
    out["ilu_momentum5"] = oracle.trifactor("ILU", s.n, rp, a.cols, a.vals)
    r = np.linspace(-1.0, 1.0, s.n)
    out["ilu_momentum5_apply"] = oracle.trifactor_apply("ILU", s.n, rp, a.cols, out["ilu_momentum5"], r)
    out["irilu_momentum5_apply"] = oracle.trifactor_apply("IRILU", s.n, rp, a.cols, out["ilu_momentum5"], r)
    H = oracle.MgHierarchy(s.n, rp, a.cols, a.vals)
    out["mg_momentum5_sizes"] = np.array([(l["n"], l["nnz"]) for l in H.levels], np.int64)
    out["mg_momentum5_agg0"] = H.levels[0]["agg"]
    out["mg_momentum5_A1_vals"] = H.levels[1]["vals"]
    out["mg_momentum5_A1_cols"] = H.levels[1]["cols"]
    out["mg_momentum5_apply"] = H.apply(r)
    s = cases.pressure_3d(5, sign=-1.0)[0]
    a, rp = csr(s)
    out["ic_pressure5"] = oracle.trifactor("IC", s.n, rp, a.cols, a.vals)
    r = np.linspace(-1.0, 1.0, s.n)
    out["ic_pressure5_apply"] = oracle.trifactor_apply("IC", s.n, rp, a.cols, out["ic_pressure5"], r)
    H = oracle.MgHierarchy(s.n, rp, a.cols, a.vals)
    out["mg_pressure5_sizes"] = np.array([(l["n"], l["nnz"]) for l in H.levels], np.int64)
    out["mg_pressure5_agg0"] = H.levels[0]["agg"]
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "precond_oracle_vectors.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
