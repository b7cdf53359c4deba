"""CPU oracle (TEST INFRASTRUCTURE ONLY): ctypes binding of oracle/oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package casadi_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_ll_p = ctypes.POINTER(ctypes.c_longlong)
c_i_p = ctypes.POINTER(ctypes.c_int)
c_d_p = ctypes.POINTER(ctypes.c_double)


def build_oracle(force=False):
    """gcc -O2 -ffp-contract=off oracle.c -> oracle/_build/liboracle.so"""
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", src,
                               "-o", so, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_oracle())
        _LIB.oracle_map_eval.restype = ctypes.c_int
        _LIB.oracle_qr_singular.restype = ctypes.c_longlong
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def map_eval(tape, N, args, want=None):
    """Serial map of an SX tape over N AoS instances (map.cpp:141-157).
    tape: dict with op,i0,i1,i2,d,sz_w,nnz_in,nnz_out.  args[j]: float64 array N*nnz_in[j] or None.
    want[j]=False leaves output j uncomputed (NULL res).  Returns list of arrays (or None)."""
    L = lib()
    n_in, n_out = len(tape["nnz_in"]), len(tape["nnz_out"])
    op = np.ascontiguousarray(tape["op"], np.int32)
    i0 = np.ascontiguousarray(tape["i0"], np.int32)
    i1 = np.ascontiguousarray(tape["i1"], np.int32)
    i2 = np.ascontiguousarray(tape["i2"], np.int32)
    d = np.ascontiguousarray(tape["d"], np.float64)
    nnz_in = np.ascontiguousarray(tape["nnz_in"], np.int64)
    nnz_out = np.ascontiguousarray(tape["nnz_out"], np.int64)
    ins = [None if a is None else np.ascontiguousarray(a, np.float64) for a in args]
    outs = []
    for j in range(n_out):
        if want is not None and not want[j]:
            outs.append(None)
        else:
            outs.append(np.full(N * int(nnz_out[j]), np.nan))
    argp = (ctypes.c_void_p * max(n_in, 1))(*[None if a is None else a.ctypes.data for a in ins])
    resp = (ctypes.c_void_p * max(n_out, 1))(*[None if a is None else a.ctypes.data for a in outs])
    w = np.zeros(int(tape["sz_w"]) + 1)
    rc = L.oracle_map_eval(ctypes.c_longlong(len(op)), _p(op, c_i_p), _p(i0, c_i_p), _p(i1, c_i_p),
                           _p(i2, c_i_p), _p(d, c_d_p), ctypes.c_longlong(n_in), _p(nnz_in, c_ll_p),
                           ctypes.c_longlong(n_out), _p(nnz_out, c_ll_p), ctypes.c_longlong(N),
                           argp, resp, _p(w, c_d_p))
    if rc != 0:
        raise RuntimeError("oracle_map_eval failed rc=%d" % rc)
    return outs


def repsum(x, nnz, n):
    L = lib()
    x = np.ascontiguousarray(x, np.float64)
    r = np.zeros(nnz)
    L.oracle_repsum(_p(x, c_d_p), ctypes.c_longlong(nnz), ctypes.c_longlong(n), _p(r, c_d_p))
    return r


def ldl_factor_solve(sp_a, sp_lt, p, a, b, nrhs=1):
    """casadi_ldl + casadi_ldl_solve for ONE system; returns (x, lt, d)."""
    L = lib()
    sp_a = np.ascontiguousarray(sp_a, np.int64); sp_lt = np.ascontiguousarray(sp_lt, np.int64)
    p = np.ascontiguousarray(p, np.int64)
    n = int(sp_lt[1])
    nnz_lt = int(sp_lt[2 + n])
    a = np.ascontiguousarray(a, np.float64)
    lt = np.zeros(nnz_lt); d = np.zeros(n); w = np.zeros(n)
    L.oracle_ldl(_p(sp_a, c_ll_p), _p(a, c_d_p), _p(sp_lt, c_ll_p), _p(lt, c_d_p), _p(d, c_d_p),
                 _p(p, c_ll_p), _p(w, c_d_p))
    x = np.array(b, np.float64, copy=True)
    L.oracle_ldl_solve(_p(x, c_d_p), ctypes.c_longlong(nrhs), _p(sp_lt, c_ll_p), _p(lt, c_d_p),
                       _p(d, c_d_p), _p(p, c_ll_p), _p(w, c_d_p))
    return x, lt, d


def qr_factor_solve(sp_a, sp_v, sp_r, prinv, pc, a, b, nrhs=1, tr=0):
    """casadi_qr + casadi_qr_solve for ONE system; returns (x, v, r, beta)."""
    L = lib()
    sp_a = np.ascontiguousarray(sp_a, np.int64); sp_v = np.ascontiguousarray(sp_v, np.int64)
    sp_r = np.ascontiguousarray(sp_r, np.int64)
    prinv = np.ascontiguousarray(prinv, np.int64); pc = np.ascontiguousarray(pc, np.int64)
    ncol = int(sp_a[1]); nrow_ext = int(sp_v[0])
    nnz_v = int(sp_v[2 + ncol]); nnz_r = int(sp_r[2 + ncol])
    a = np.ascontiguousarray(a, np.float64)
    v = np.zeros(nnz_v); r = np.zeros(nnz_r); beta = np.zeros(ncol)
    w = np.zeros(max(nrow_ext, ncol) + ncol)
    L.oracle_qr(_p(sp_a, c_ll_p), _p(a, c_d_p), _p(w, c_d_p), _p(sp_v, c_ll_p), _p(v, c_d_p),
                _p(sp_r, c_ll_p), _p(r, c_d_p), _p(beta, c_d_p), _p(prinv, c_ll_p), _p(pc, c_ll_p))
    x = np.array(b, np.float64, copy=True)
    L.oracle_qr_solve(_p(x, c_d_p), ctypes.c_longlong(nrhs), ctypes.c_int(tr), _p(sp_v, c_ll_p),
                      _p(v, c_d_p), _p(sp_r, c_ll_p), _p(r, c_d_p), _p(beta, c_d_p), _p(prinv, c_ll_p),
                      _p(pc, c_ll_p), _p(w, c_d_p))
    return x, v, r, beta


def mtimes(x, sp_x, y, sp_y, z, sp_z):
    L = lib()
    sp_x = np.ascontiguousarray(sp_x, np.int64); sp_y = np.ascontiguousarray(sp_y, np.int64)
    sp_z = np.ascontiguousarray(sp_z, np.int64)
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    z = np.array(z, np.float64, copy=True)
    w = np.zeros(int(sp_x[0]))
    L.oracle_mtimes(_p(x, c_d_p), _p(sp_x, c_ll_p), _p(y, c_d_p), _p(sp_y, c_ll_p), _p(z, c_d_p),
                    _p(sp_z, c_ll_p), _p(w, c_d_p))
    return z


def vec_op(op, x, y):
    """Elementwise reference scalar op (opcode numbering of calculus.hpp) on float64 arrays."""
    L = lib()
    x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
    f = np.empty_like(x)
    rc = L.oracle_vec_op(ctypes.c_int(op), ctypes.c_longlong(x.size), _p(x, c_d_p), _p(y, c_d_p), _p(f, c_d_p))
    if rc:
        raise RuntimeError("oracle_vec_op: opcode %d not evaluable" % op)
    return f
