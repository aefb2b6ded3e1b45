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


def _function_sass(obj, name_part):
    """SASS text of the kernels of one object file of the build whose mangled name contains `name_part`."""
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        pytest.skip(f"{obj} not built in-tree")
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    return [chunk for chunk in text.split("Function : ")[1:] if name_part in chunk.split("\n", 1)[0]]


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_of_the_sweep_and_merge_path_kernels(table):
    """The dependency-driven triangular sweeps poll and publish with gpu-scope relaxed accesses and vote on
    the warp's completion inside ONE loop (no lane ever parks outside it while another still polls); the
    merge-path SpMV reads its slice with 128-bit loads and parks the products in shared memory."""
    sweeps = _function_sass("trifactor.o", "k_trisolve_sf_short")
    assert len(sweeps) == 6                                     # {lower unit, lower, upper} x {4, 8 operands}
    for sass in sweeps:
        assert "LDG.E.64.STRONG.GPU" in sass and "STG.E.64.STRONG.GPU" in sass
        assert "VOTE.ALL" in sass
        # the publishing store sits before the vote in program order (it is issued from inside the loop)
        assert sass.index("STG.E.64.STRONG.GPU") < sass.index("VOTE.ALL")
    merge = _function_sass("spmv_merge.o", "k_spmv_merge")
    main = [s for s in merge if "fixup" not in s.split("\n", 1)[0] and "_dot" not in s.split("\n", 1)[0]]
    assert len(main) == 2
    for sass in main:
        assert re.search(r"LDG\.E(\.EF)?\.128", sass) and "STS.128" in sass and "LDS" in sass
    # co-residency budget of the sweeps: the 8-operand kernels keep 2 CTAs of 256 threads per SM
    for parts in (("k_trisolve_sf_short<true, true, 8>",), ("k_trisolve_sf_short<false, false, 8>",)):
        regs, _, _ = find(table, *parts)
        assert regs <= 128
