"""Merge-path SpMV (spmv_variant 8) against the row-based kernels on matrices with one / a few very long
rows, and on the regular 100^3 mesh matrix (where it must not be picked).  python tools/merge_probe.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ogl_b200 import cases  # noqa: E402
from ogl_b200.backend import Context  # noqa: E402


def hub_system(n, hubs, seed=1):
    """`hubs` cells coupled to every other cell + a chain: `hubs` rows of ~n entries among rows of ~hubs + 3."""
    lo, up = [], []
    for h in range(hubs):
        lo.append(np.full(n - h - 1, h, np.int32))
        up.append(np.arange(h + 1, n, dtype=np.int32))
    lo.append(np.arange(hubs, n - 1, dtype=np.int32))
    up.append(np.arange(hubs + 1, n, dtype=np.int32))
    lower, upper = np.concatenate(lo), np.concatenate(up)
    key = np.unique(lower.astype(np.int64) * n + upper)
    lower, upper = (key // n).astype(np.int32), (key % n).astype(np.int32)
    rng = np.random.default_rng(seed)
    return n, lower, upper, rng.uniform(1, 2, n) * n, rng.normal(size=lower.size)


for name, (n, lower, upper, diag, up) in (("1 hub, 2 M rows", hub_system(2000000, 1)),
                                          ("8 hubs, 1 M rows", hub_system(1000000, 8))):
    ctx = Context()
    ctx.pattern_from_ldu(n, lower, upper, True)
    ctx.values_update(diag, up)
    auto = ctx.get_option("spmv_variant_in_use")
    for variant in (8, 3, 2):
        ctx.set_option("spmv_variant", variant)
        for fused in (False, True):
            ctx.spmv_bench(3, fused_dot=fused)
            ms = ctx.spmv_bench(20, fused_dot=fused) / 20
            print(json.dumps({"matrix": name, "rows": n, "nnz": int(ctx.nnz), "auto_variant": auto, "variant": variant,
                              "fused_dot": fused, "us": round(ms * 1e3, 1),
                              "GBs_alg": round((12 * ctx.nnz + 20 * n) / ms / 1e6, 1)}), flush=True)
    ctx.close()
s = cases.pressure_3d(100)[0]
ctx = Context()
ctx.pattern_from_ldu(s.n, s.lower_addr, s.upper_addr, True)
ctx.values_update(s.diag, s.upper)
auto = ctx.get_option("spmv_variant_in_use")
for variant in (8, 6, 7):
    ctx.set_option("spmv_variant", variant)
    ctx.spmv_bench(3, fused_dot=False)
    ms = ctx.spmv_bench(50, fused_dot=False) / 50
    print(json.dumps({"matrix": "100^3 pressure", "rows": s.n, "nnz": int(ctx.nnz), "auto_variant": auto,
                      "variant": variant, "fused_dot": False, "us": round(ms * 1e3, 1),
                      "GBs_alg": round((12 * ctx.nnz + 20 * s.n) / ms / 1e6, 1)}), flush=True)
ctx.close()
