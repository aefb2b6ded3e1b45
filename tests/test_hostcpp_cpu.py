"""CPU: the C++ plugin layer loads, its run-time selection table and keyword
checks behave like the reference's, and it fails loudly without a GPU."""
import pytest

from foam_harness import FoamCase, FoamFatalError, dict_text
from ogl_b200 import cases

BASE = {"solver": "GKOCG", "preconditioner": "BJ", "executor": "cuda", "tolerance": 1e-8,
        "relTol": 0.0, "adaptMinIter": False}


def test_dictionary_text():
    t = dict_text({"solver": "GKOCG", "relTol": 0.0, "adaptMinIter": False,
                   "preconditioner": {"preconditioner": "BJ", "maxBlockSize": 4}})
    assert "solver GKOCG;" in t and "adaptMinIter false;" in t
    assert "preconditioner { preconditioner BJ; maxBlockSize 4; }" in t


def test_selection_table_and_keyword_errors():
    sym = FoamCase(cases.pressure_3d(4)[0])
    asym = FoamCase(cases.momentum_3d(4)[0])
    try:
        s = sym.s
        # GKOCG is registered for symmetric matrices only (Solver/CG/GKOCG.C:16-17)
        with pytest.raises(FoamFatalError, match="Unknown asymmetric matrix solver GKOCG"):
            asym.solve("U", BASE, asym.s.psi, asym.s.source)
        with pytest.raises(FoamFatalError, match="Unknown symmetric matrix solver PCG"):
            sym.solve("p", dict(BASE, solver="PCG"), s.psi, s.source)
        with pytest.raises(FoamFatalError, match="unknown matrixFormat Hybrid"):
            sym.solve("p", dict(BASE, matrixFormat="Hybrid"), s.psi, s.source)
        with pytest.raises(FoamFatalError, match="does not support the executor: reference"):
            sym.solve("p", {k: v for k, v in BASE.items() if k != "executor"}, s.psi, s.source)
    finally:
        sym.close()
        asym.close()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = FoamCase(cases.pressure_3d(4)[0])
    try:
        with pytest.raises(FoamFatalError, match="cannot create the CUDA executor"):
            c.solve("p", BASE, c.s.psi, c.s.source)
    finally:
        c.close()
