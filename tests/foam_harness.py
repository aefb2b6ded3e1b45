"""ctypes driver of ogl_b200/libOGL_b200.so (the C++ plugin layer built against
the OpenFOAM shim): builds an lduMatrix + interfaces from an LduSystem and runs
lduMatrix::solver::New(...)->solve(...) with an fvSolution-style dictionary."""
import ctypes as C
import os

import numpy as np

from ogl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "ogl_b200", "libOGL_b200.so")
_h = None


def lib():
    global _h
    if _h is None:
        _lib.load()   # maps NCCL + libogl_b200.so first
        h = C.CDLL(PATH)
        h.foamshim_last_error.restype = C.c_char_p
        h.foamshim_case_create.restype = C.c_void_p
        h.foamshim_case_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        h.foamshim_case_destroy.argtypes = [C.c_void_p]
        h.foamshim_registry_size.argtypes = [C.c_void_p]
        h.foamshim_case_set_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        h.foamshim_solve.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_int)]
        h.foamshim_set_parallel.argtypes = [C.c_int, C.c_int, C.c_char_p]
        _h = h
    return _h


class FoamFatalError(RuntimeError):
    pass


def dict_text(controls: dict) -> str:
    out = []
    for k, v in controls.items():
        if isinstance(v, dict):
            out.append(f"{k} {{ {dict_text(v)} }}")
        elif isinstance(v, bool):
            out.append(f"{k} {'true' if v else 'false'};")
        else:
            out.append(f"{k} {v!r};" if isinstance(v, float) else f"{k} {v};")
    return " ".join(out)


class FoamCase:
    KIND = {"processor": 0, "cyclic": 1, "cyclicAMI": 2}

    def __init__(self, s):
        self.s = s
        kinds = np.array([self.KIND[i.kind] for i in s.interfaces], np.int32)
        nbr = np.array([i.nbr_rank if i.kind == "processor" else i.nbr_patch for i in s.interfaces],
                       np.int32)
        sizes = np.array([i.face_cells.size for i in s.interfaces], np.int32)
        fcs = (np.concatenate([i.face_cells for i in s.interfaces]).astype(np.int32)
               if s.interfaces else np.zeros(1, np.int32))
        lo, up = (np.ascontiguousarray(a, np.int32) for a in (s.lower_addr, s.upper_addr))
        self.h = lib().foamshim_case_create(s.n, lo.size, lo.ctypes.data, up.ctypes.data,
                                            len(s.interfaces), kinds.ctypes.data, nbr.ctypes.data,
                                            sizes.ctypes.data, fcs.ctypes.data)
        self.set_coeffs(s)

    def set_coeffs(self, s):
        bou = (np.concatenate([i.bou_coeffs for i in s.interfaces]).astype(np.float64)
               if s.interfaces else np.zeros(1))
        d, u = (np.ascontiguousarray(a, np.float64) for a in (s.diag, s.upper))
        lw = None if s.symmetric else np.ascontiguousarray(s.lower, np.float64)
        lib().foamshim_case_set_coeffs(self.h, s.n, u.size, d.ctypes.data, u.ctypes.data,
                                       lw.ctypes.data if lw is not None else None, bou.ctypes.data)

    def registry_size(self):
        return lib().foamshim_registry_size(self.h)

    def solve(self, field, controls, psi, source):
        psi = np.array(psi, dtype=np.float64)   # copy: the caller's psi stays untouched
        src = np.ascontiguousarray(source, np.float64)
        name = C.create_string_buffer(128)
        a, b, it = C.c_double(0), C.c_double(0), C.c_int(0)
        rc = lib().foamshim_solve(self.h, field.encode(), dict_text(controls).encode(), self.s.n,
                                  psi.ctypes.data, src.ctypes.data, name, 128, C.byref(a),
                                  C.byref(b), C.byref(it))
        if rc:
            raise FoamFatalError(lib().foamshim_last_error().decode())
        return psi, name.value.decode(), a.value, b.value, it.value

    def close(self):
        if self.h:
            lib().foamshim_case_destroy(self.h)
            self.h = None
