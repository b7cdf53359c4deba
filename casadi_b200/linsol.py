"""Batched linear solves with sparsity shared across the batch (ccu_ldl_create / ccu_qr_create).

Python mirror of the numeric phase of the reference's Linsol plugins "ldl" and "qr"
(casadi/solvers/linsol_ldl.cpp:119-132, linsol_qr.cpp:126-180); the symbolic phase (Sparsity::ldl,
Sparsity::qr_sparse) stays on the host in the reference and is passed in as compressed CCS vectors.
"""
import ctypes

import numpy as np

from . import capi
from .capi import CcuError
from .cuda_map import CudaTape


def _ll(a):
    a = np.ascontiguousarray(a, np.int64)
    return a, a.ctypes.data_as(capi.c_ll_p)


class _BorrowedTape(CudaTape):
    """View of a tape owned by another object (never destroyed from here)."""

    def __init__(self, handle, nnz_in, nnz_out, device):
        self.handle, self.nnz_in, self.nnz_out, self.device = handle, list(nnz_in), list(nnz_out), device

    def close(self):
        self.handle = None


class CudaLinsol:
    def __init__(self, kind, sp_a, sym, nrhs=1, tr=False, eps=1e-12, device=0, mode=None):
        """kind "ldl": sym = (sp_lt, p);  kind "qr": sym = (sp_v, sp_r, prinv, pc)."""
        L = capi.lib()
        self._keep = []
        sp_a, p_a = _ll(sp_a)
        self.n, self.nnz_a, self.nrhs, self.kind = int(sp_a[1]), int(sp_a[2 + int(sp_a[1])]), int(nrhs), kind
        ptrs = []
        for s in sym:
            arr, ptr = _ll(s)
            self._keep.append(arr)
            ptrs.append(ptr)
        capi.check(L.ccu_set_default_mode(capi.MODES[mode]))
        try:
            if kind == "ldl":
                self.handle = L.ccu_ldl_create(p_a, ptrs[0], ptrs[1], self.nrhs, int(device))
            elif kind == "qr":
                self.handle = L.ccu_qr_create(p_a, ptrs[0], ptrs[1], ptrs[2], ptrs[3], self.nrhs, int(bool(tr)), float(eps),
                                              int(device))
            else:
                raise CcuError("unknown linear solver kind %r" % kind)
        finally:
            L.ccu_set_default_mode(-1)
        if not self.handle:
            raise CcuError(capi.last_error())
        self.tape = _BorrowedTape(L.ccu_linsol_tape(self.handle), [self.nnz_a, self.n * self.nrhs], [self.n * self.nrhs, 1], device)

    def close(self):
        if getattr(self, "handle", None):
            capi.lib().ccu_linsol_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, A, B):
        """A: (N, nnz(A)) values in CCS order, B: (N, n*nrhs).  Returns (X, number of flagged instances)."""
        A = np.ascontiguousarray(A, np.float64)
        B = np.ascontiguousarray(B, np.float64)
        N = A.size // self.nnz_a
        if A.size != N * self.nnz_a or B.size != N * self.n * self.nrhs:
            raise CcuError("inconsistent batch sizes")
        X = np.full(B.shape, np.nan)
        flagged = ctypes.c_longlong(0)
        capi.check(capi.lib().ccu_linsol_solve_host(self.handle, N, A.ctypes.data_as(capi.c_d_p), B.ctypes.data_as(capi.c_d_p),
                                                    X.ctypes.data_as(capi.c_d_p), ctypes.byref(flagged)))
        return X, int(flagged.value)
