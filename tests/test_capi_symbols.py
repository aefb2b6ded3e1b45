"""CPU: the C-ABI library loads and exports every symbol include/ogl_b200.h
declares; without a GPU every compute entry point fails loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ogl_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ogl_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("ogl_ctx_create", "ogl_pattern_from_ldu", "ogl_partition_create",
                 "ogl_nonlocal_pattern", "ogl_values_update", "ogl_vector_upload",
                 "ogl_precond_setup", "ogl_solve", "ogl_spmv", "ogl_export_mtx"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from ogl_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True).stdout
    exported = set(re.findall(r" T (ogl_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_library_is_sm100a_only():
    from ogl_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ogl_b200.backend import Context, OglError
    with pytest.raises(OglError) as e:
        Context()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ogl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".H", ".C")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "oracle/" not in txt, f


def test_preconditioner_keywords_match_the_header_enum():
    """The `preconditioner` keyword table of the Python host layer, the ctypes constants, the C++ host
    layer and the oracle agree with include/ogl_b200.h (same names, same numbers)."""
    from ogl_b200 import _lib
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    enum = dict((k, int(v)) for k, v in re.findall(r"OGL_PRECOND_([A-Z]+)\s*=\s*(\d+)", txt))
    assert enum == {"NONE": 0, "BJ": 1, "ISAI": 2, "GISAI": 3, "ILU": 4, "IC": 5, "IRILU": 6, "MULTIGRID": 7}
    for name, value in enum.items():
        assert getattr(_lib, "OGL_PRECOND_" + name) == value
    src = open(os.path.join(ROOT, "ogl_b200", "plugin.py")).read()
    table = src[src.index("PRECOND_KINDS = {"):src.index("}", src.index("PRECOND_KINDS = {"))]
    words = dict(re.findall(r'"(\w+)":\s*L\.OGL_PRECOND_(\w+)', table))
    assert {w.upper(): k for w, k in words.items()} == {k: k for k in enum}
    cpp = open(os.path.join(ROOT, "ogl_b200", "host_cpp", "Preconditioner.H")).read()
    for word, kind in words.items():
        assert re.search(r'name == "%s"\) kind = OGL_PRECOND_%s;' % (word, kind), cpp), word
    import oracle
    assert {k.upper(): v for k, v in oracle.PRECONDS.items()} == enum
