"""Generates tests/golden/ldu_ref_vectors.npz with the REFERENCE's own free
functions (oracle/_ref/libogl_ref.so, compiled from
/root/reference/HostMatrix/HostMatrixFreeFunctions.C) on seeded random meshes.
Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The three known-answer vectors of /root/reference/unitTests/test_HostMatrix.C
(:8-107) are restated in tests/test_oracle_golden.py itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from conftest import random_ldu_mesh  # noqa: E402


def main():
    assert oracle.ref_lib() is not None, "oracle/_ref not built (needs /root/reference)"
    rng = np.random.default_rng(20240621)
    out = {}
    cases = [(5, 3), (17, 20), (64, 150), (200, 900), (1000, 2500)]
    for k, (n, extra) in enumerate(cases):
        lower, upper = random_ldu_mesh(rng, n, extra)
        F = lower.size
        for sym in (True, False):
            tag = f"c{k}_{'sym' if sym else 'asym'}"
            rows, cols, perm = oracle.init_local_sparsity(n, upper, lower, sym, which="ref")
            diag = rng.uniform(1, 2, n)
            up = rng.uniform(-1, 0, F)
            lo = rng.uniform(-1, 0, F)
            scale = 1.0 if k % 2 == 0 else -2.5
            if sym:
                vals = oracle.update_host("symmetric", perm, scale, diag, up, which="ref")
            else:
                vals = oracle.update_host("non_symmetric", perm, scale, diag, up, lo, which="ref")
            out.update({f"{tag}_n": n, f"{tag}_lower": lower, f"{tag}_upper": upper,
                        f"{tag}_rows": rows, f"{tag}_cols": cols, f"{tag}_perm": perm,
                        f"{tag}_diag": diag, f"{tag}_up": up, f"{tag}_lo": lo,
                        f"{tag}_scale": scale, f"{tag}_vals": vals})
    np.savez_compressed(os.path.join(HERE, "ldu_ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
