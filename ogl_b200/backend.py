"""Thin object wrapper over the C ABI: one `Context` per (device, field)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


class OglError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ogl_b200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    """numpy array / torch tensor / int address / None -> void*."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    if a is None or isinstance(a, int) or hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(L.OGL_NCCL_ID_BYTES)
    rc = L.load().ogl_nccl_unique_id(buf)
    if rc:
        raise OglError(rc, (L.load().ogl_last_error(None) or b"").decode())
    return buf.raw


def device_count() -> int:
    n = C.c_int(0)
    L.load().ogl_device_count(C.byref(n))
    return n.value


class Context:
    def __init__(self, device_id: int = 0, rank: int = 0, n_ranks: int = 1,
                 nccl_id: Optional[bytes] = None, stream: int = 0):
        self.lib = L.load()
        self.h = L.ctx_p()
        idbuf = C.create_string_buffer(nccl_id, L.OGL_NCCL_ID_BYTES) if nccl_id else None
        rc = self.lib.ogl_ctx_create(device_id, rank, n_ranks, idbuf,
                                     C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc:
            raise OglError(rc, (self.lib.ogl_last_error(None) or b"").decode())
        self.rank, self.n_ranks = rank, n_ranks
        self.n = 0
        self.nnz = 0
        self.n_halo = 0

    def _check(self, rc):
        if rc:
            raise OglError(rc, (self.lib.ogl_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.lib.ogl_ctx_destroy(self.h)
            self.h = L.ctx_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options
    def set_option(self, key: str, value: int):
        self._check(self.lib.ogl_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64(0)
        self._check(self.lib.ogl_get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    # -- pattern
    def pattern_from_ldu(self, n, lower_addr, upper_addr, symmetric, iface_rows=None,
                         iface_cols=None):
        lower_addr, upper_addr = _i32(lower_addr), _i32(upper_addr)
        ir = _i32(iface_rows if iface_rows is not None else [])
        ic = _i32(iface_cols if iface_cols is not None else [])
        self._check(self.lib.ogl_pattern_from_ldu(
            self.h, n, lower_addr.size, int(bool(symmetric)), _ptr(lower_addr), _ptr(upper_addr),
            ir.size, _ptr(ir), _ptr(ic)))
        self.n = int(n)
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ogl_pattern_nnz(self.h, C.byref(a), C.byref(b)))
        self.nnz, self.n_halo = a.value, 0

    def pattern_download(self):
        rows, cols, mp = (np.empty(self.nnz, np.int32) for _ in range(3))
        rp = np.empty(self.n + 1, np.int32)
        self._check(self.lib.ogl_pattern_download(self.h, _ptr(rows), _ptr(cols), _ptr(mp), _ptr(rp)))
        return rows, cols, mp, rp

    def partition_create(self, n_local, target_ids, target_sizes, send_idxs):
        t, s, i = _i32(target_ids), _i32(target_sizes), _i32(send_idxs)
        self._check(self.lib.ogl_partition_create(self.h, n_local, t.size, _ptr(t), _ptr(s), _ptr(i)))

    def partition_export(self) -> bytes:
        """This rank's window directory (host-driven bootstrap, contexts without an NCCL id)."""
        size = C.c_int64(0)
        self._check(self.lib.ogl_partition_export(self.h, None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        self._check(self.lib.ogl_partition_export(self.h, buf, size.value, C.byref(size)))
        return buf.raw

    def partition_connect(self, blobs):
        """`blobs`: every rank's directory in rank order (after an all-gather by the caller)."""
        raw = b"".join(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        self._check(self.lib.ogl_partition_connect(self.h, buf, len(blobs)))

    def partition_sizes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ogl_partition_sizes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def nonlocal_pattern(self, face_cells):
        fc = _i32(face_cells)
        self._check(self.lib.ogl_nonlocal_pattern(self.h, fc.size, _ptr(fc)))
        self.n_halo = int(fc.size)

    def nonlocal_pattern_download(self):
        r, c, m = (np.empty(self.n_halo, np.int32) for _ in range(3))
        self._check(self.lib.ogl_nonlocal_pattern_download(self.h, _ptr(r), _ptr(c), _ptr(m)))
        return r, c, m

    # -- values / vectors (numpy arrays, pinned torch tensors or raw addresses)
    def values_update(self, diag, upper, lower=None, local_iface_bou=None, nonlocal_bou=None,
                      scaling=1.0):
        keep = [_f64(a) for a in (diag, upper, lower, local_iface_bou, nonlocal_bou)]
        self._check(self.lib.ogl_values_update(self.h, *[_ptr(a) for a in keep], float(scaling)))

    def values_download(self):
        v, nl = np.empty(self.nnz), np.empty(max(self.n_halo, 1))
        self._check(self.lib.ogl_values_download(self.h, _ptr(v), _ptr(nl)))
        return v, nl[:self.n_halo]

    def vector_upload(self, which, host, scale=1.0):
        host = _f64(host)
        self._check(self.lib.ogl_vector_upload(self.h, which, _ptr(host), float(scale)))

    def vector_download(self, which, out=None):
        if out is None:
            out = np.empty(self.n)
        self._check(self.lib.ogl_vector_download(self.h, which, _ptr(out)))
        return out

    def vector_fill(self, which, value):
        self._check(self.lib.ogl_vector_fill(self.h, which, float(value)))

    # -- preconditioner / solve
    def precond_setup(self, kind, max_block_size=1, skip_sorting=True):
        self._check(self.lib.ogl_precond_setup(self.h, kind, max_block_size, int(skip_sorting)))

    def precond_download(self):
        nb = C.c_int32(0)
        self._check(self.lib.ogl_precond_download(self.h, C.byref(nb), None, None))
        bp = np.empty(nb.value + 1, np.int32)
        self._check(self.lib.ogl_precond_download(self.h, C.byref(nb), _ptr(bp), None))
        inv = np.empty(int((np.diff(bp).astype(np.int64) ** 2).sum()))
        self._check(self.lib.ogl_precond_download(self.h, C.byref(nb), _ptr(bp), _ptr(inv)))
        return bp, inv

    def precond_factors_download(self):
        """ILU / IC / IRILU factors over the local CSR pattern (parity hook)."""
        out = np.empty(self.nnz)
        self._check(self.lib.ogl_precond_factors_download(self.h, _ptr(out)))
        return out

    def mg_levels(self):
        """The Multigrid hierarchy, level by level (parity hook): dicts with n, nnz, n_coarse, row_ptrs,
        cols, vals, agg (None on the coarsest level)."""
        nl = C.c_int32(0)
        self._check(self.lib.ogl_mg_levels(self.h, C.byref(nl)))
        out = []
        for l in range(nl.value):
            n, nnz, nc = C.c_int32(0), C.c_int32(0), C.c_int32(0)
            self._check(self.lib.ogl_mg_level_info(self.h, l, C.byref(n), C.byref(nnz), C.byref(nc)))
            rp, cols = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32)
            vals = np.empty(nnz.value)
            agg = np.empty(n.value, np.int32) if nc.value > 0 else None
            self._check(self.lib.ogl_mg_level_download(self.h, l, _ptr(rp), _ptr(cols), _ptr(vals),
                                                       _ptr(agg) if agg is not None else None))
            out.append(dict(n=n.value, nnz=nnz.value, n_coarse=nc.value, row_ptrs=rp, cols=cols, vals=vals, agg=agg))
        return out

    def precond_apply(self, r):
        """z = M^-1 r on host vectors of the local size (parity hook)."""
        r = np.ascontiguousarray(r, np.float64)
        z = np.empty_like(r)
        self._check(self.lib.ogl_precond_apply(self.h, _ptr(r), _ptr(z)))
        return z

    def solve(self, solver, tolerance=1e-6, rel_tol=0.0, min_iter=0, max_iter=1000, frequency=1,
              krylov_dim=100, export_res=False) -> L.SolveResult:
        p = L.SolveParams(solver, tolerance, rel_tol, min_iter, max_iter, frequency, krylov_dim,
                          int(export_res))
        r = L.SolveResult()
        self._check(self.lib.ogl_solve(self.h, C.byref(p), C.byref(r)))
        return r

    def residual_history(self, capacity):
        out = np.zeros(capacity)
        n = C.c_int32(0)
        self._check(self.lib.ogl_residual_history(self.h, _ptr(out), capacity, C.byref(n)))
        return out[:n.value]

    def spmv(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        self._check(self.lib.ogl_spmv(self.h, _ptr(x), _ptr(y)))
        return y

    def spmv_bench(self, reps, fused_dot=False) -> float:
        ms = C.c_float(0)
        self._check(self.lib.ogl_spmv_bench(self.h, reps, int(fused_dot), C.byref(ms)))
        return ms.value

    def pcg_bench(self, iters) -> float:
        ms = C.c_float(0)
        self._check(self.lib.ogl_pcg_bench(self.h, iters, C.byref(ms)))
        return ms.value

    def trace_download(self, cap: int = 1 << 16):
        """Timeline events logged since the last call (option `trace` 1): (tags, times_ns)."""
        ev = np.zeros(cap, dtype=np.uint64)
        n = C.c_int64(0)
        self._check(self.lib.ogl_trace_download(self.h, ev.ctypes.data, cap, C.byref(n)))
        ev = ev[: n.value]
        return (ev >> np.uint64(48)).astype(np.int64), (ev & np.uint64((1 << 48) - 1)).astype(np.int64)

    def membench(self, mode: int, n_doubles: int = 1 << 27, reps: int = 10) -> float:
        g = C.c_double(0)
        self._check(self.lib.ogl_membench(self.h, mode, n_doubles, reps, C.byref(g)))
        return g.value

    def commbench(self, mode: int, reps: int = 200) -> float:
        u = C.c_double(0)
        self._check(self.lib.ogl_commbench(self.h, mode, reps, C.byref(u)))
        return u.value

    def synchronize(self):
        self._check(self.lib.ogl_synchronize(self.h))

    def export_mtx(self, which, path):
        self._check(self.lib.ogl_export_mtx(self.h, which, str(path).encode()))
