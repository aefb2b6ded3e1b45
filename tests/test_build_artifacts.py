"""CPU-side guards on what nvcc/ptxas produced for the hot kernels (no GPU needed): register
budgets that the occupancy of the persistent grids relies on, no local-memory spills, and the
instructions that prove the TMA / graph-conditional / acq_rel paths are really in the binary.
Skipped when the library has not been built in-tree (`__graft_entry__.build()`)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ogl_b200", "csrc", "build")
LIB = os.path.join(ROOT, "ogl_b200", "libogl_b200.so")


def ptxas_table():
    """demangled kernel name -> (registers, spill store bytes, spill load bytes)."""
    out = {}
    pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes "
                     r"spill stores, (\d+) bytes spill loads.*?Used (\d+) registers", re.S)
    for name in os.listdir(BUILD):
        if not name.endswith(".ptxas.log"):
            continue
        text = open(os.path.join(BUILD, name)).read()
        for m in pat.finditer(text):
            out[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    if not out:
        pytest.skip("no ptxas logs")
    mangled = list(out)
    dem = subprocess.run(["c++filt"] + mangled, capture_output=True, text=True).stdout.splitlines()
    return {d: out[m] for d, m in zip(dem, mangled)}


@pytest.fixture(scope="module")
def table():
    if not os.path.isdir(BUILD) or shutil.which("c++filt") is None:
        pytest.skip("library not built in-tree")
    return ptxas_table()


def find(table, *parts):
    hits = [(k, v) for k, v in table.items() if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, [k for k, _ in hits])
    return hits[0][1]


def test_hot_kernels_fit_their_occupancy_budget(table):
    # 5 CTAs x 256 threads per SM: <= 51 registers; the CG loop's instantiations must not spill
    for parts in (("k_spmv_pipe<false, 1, false>",), ("k_spmv_pipe<false, 0, false>",),
                  ("k_spmv_pipe<true, 0, false>",), ("k_pcg_fused<1>",), ("k_pcg_fused<0>",)):
        regs, st, ld = find(table, *parts)
        assert regs <= 51 and st == 0 and ld == 0, (parts, regs, st, ld)
    # 4 CTAs x 256 threads per SM: <= 64 registers
    for parts in (("k_cg_xr<1>",), ("k_cg_xr<0>",), ("k_cg_p(",), ("k_spmv_ell<false, 1, 7, true, 4>",),
                  ("k_spmv_ell<false, 0, 7, true, 4>",), ("k_spmv_ell_cgp<7, false, false, 4>",)):
        regs, st, ld = find(table, *parts)
        assert regs <= 64 and st == 0 and ld == 0, (parts, regs, st, ld)


def test_only_sm100a_code(table):
    logs = "".join(open(os.path.join(BUILD, f)).read() for f in os.listdir(BUILD) if f.endswith(".ptxas.log"))
    assert "sm_100a" in logs
    assert not re.search(r"for 'sm_(?!100a)", logs)


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_carries_the_paths_the_design_claims():
    if not os.path.exists(LIB):
        pytest.skip("library not built in-tree")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass                      # TMA bulk copies of the k_spmv_tma pipeline
    assert "SYNCS" in sass                       # its mbarriers
    assert "ATOMG.E.ADD.STRONG.GPU" in sass      # acq_rel tickets of the reductions / grid barrier
    assert "LDG.E.128" in sass and "STG.E.128" in sass   # 128-bit vector updates
    assert "HMMA" not in sass and "UTCHMMA" not in sass  # HBM-bound FP64 path: no tensor-core reshaping
