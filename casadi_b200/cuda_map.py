"""Python mirror of the reference's Map interface for the "cuda" parallelization.

`CudaTape` wraps an exported SXFunction tape (the `f` of `f.map(N, "cuda")`); `CudaMap` mirrors
casadi::Map (casadi/core/map.hpp:50-170): same argument meaning (AoS host buffers, None = NULL
argument/result) and error behaviour (exceptions carry the library's message).  All computation is
in libcasadi_cuda.so; this file only marshals pointers.
"""
import ctypes

import numpy as np

from . import capi
from .capi import CcuError, LAYOUT_AOS, LAYOUT_SOA  # noqa: F401


class CudaTape:
    """A compiled SX tape resident on one device (ccu_tape)."""

    def __init__(self, tape, device=0, mode=None):
        """mode: None (environment variable CCU_MODE, default auto), "interp", "jit" or "auto"."""
        L = capi.lib()
        self.nnz_in = [int(v) for v in tape["nnz_in"]]
        self.nnz_out = [int(v) for v in tape["nnz_out"]]
        op = np.ascontiguousarray(tape["op"], np.int32)
        i0 = np.ascontiguousarray(tape["i0"], np.int32)
        i1 = np.ascontiguousarray(tape["i1"], np.int32)
        i2 = np.ascontiguousarray(tape["i2"], np.int32)
        d = np.ascontiguousarray(tape["d"], np.float64)
        nin = np.ascontiguousarray(self.nnz_in, np.int64)
        nout = np.ascontiguousarray(self.nnz_out, np.int64)
        p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
        capi.check(L.ccu_set_default_mode(capi.MODES[mode]))
        try:
            self.handle = L.ccu_tape_create(len(op), p(op, capi.c_i_p), p(i0, capi.c_i_p), p(i1, capi.c_i_p),
                                            p(i2, capi.c_i_p), p(d, capi.c_d_p), int(tape["sz_w"]), len(nin),
                                            p(nin, capi.c_ll_p), len(nout), p(nout, capi.c_ll_p), int(device))
        finally:
            L.ccu_set_default_mode(-1)
        if not self.handle:
            raise CcuError(capi.last_error())
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            capi.lib().ccu_tape_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_in(self):
        return len(self.nnz_in)

    @property
    def n_out(self):
        return len(self.nnz_out)

    def info(self):
        inf = capi.TapeInfo()
        capi.check(capi.lib().ccu_tape_get_info(self.handle, ctypes.byref(inf)))
        return {n: int(getattr(inf, n)) for n, _ in inf._fields_}

    def set_plan(self, threads=0, ipt=0, slots_shared=0):
        capi.check(capi.lib().ccu_tape_set_plan(self.handle, threads, ipt, slots_shared))

    def set_mode(self, mode):
        """capi.MODE_INTERP or capi.MODE_JIT (raises when the tape cannot be specialised)."""
        capi.check(capi.lib().ccu_tape_set_mode(self.handle, mode))

    def set_jit_plan(self, seg_instr=0, threads=0, min_blocks=-1, tile=-1):
        capi.check(capi.lib().ccu_tape_set_jit_plan(self.handle, seg_instr, threads, min_blocks, tile))

    def set_jit_schedule(self, schedule):
        """0 = reference tape order, 1 = min-cut bisection order (csrc/tape_schedule.hpp); takes effect at the
        next (re)build of the specialised kernels."""
        capi.check(capi.lib().ccu_tape_set_jit_schedule(self.handle, schedule))

    def set_jit_remat(self, remat):
        """Rematerialisation price (FP64 issue slots per cross-segment value; 0 = off); rebuilds the kernels."""
        capi.check(capi.lib().ccu_tape_set_jit_remat(self.handle, remat))

    def jit_remat_stats(self, seg_instr=0, remat=-1):
        st = (ctypes.c_longlong * 6)()
        capi.check(capi.lib().ccu_tape_jit_remat_stats(self.handle, seg_instr, remat, st))
        return dict(cloned=st[0], dropped=st[1], cross_loads=st[2], cross_stores=st[3], segments=st[4], scratch_slots=st[5])

    def loop_stats(self):
        """The time-stepping loop recovered from the unrolled tape (csrc/tape_reroll.hpp), host only."""
        st = (ctypes.c_longlong * 8)()
        capi.check(capi.lib().ccu_tape_loop_stats(self.handle, st))
        d = dict(found=bool(st[0]), iters=st[1], body=st[2], carried=st[3], varying_constants=st[4], advancing_inputs=st[5],
                 exits=st[6], before=st[7])
        if not d["found"]:
            d["why"] = capi.last_error()
        return d

    def jit_plan_stats(self, seg_instr=0, schedule=-1):
        """Plan of the specialisation (host only, nothing compiled)."""
        st = (ctypes.c_longlong * 8)()
        capi.check(capi.lib().ccu_tape_jit_plan_stats(self.handle, seg_instr, schedule, st))
        return dict(segments=st[0], scratch_slots=st[1], cross_loads=st[2], cross_stores=st[3], max_segment=st[4],
                    schedule_ms=st[5], max_live=st[6], mean_live=st[7])

    def jit_link_check(self):
        """Size of the linked persistent-chain cubin of the current plan (host only); raises when NVRTC / nvJitLink
        cannot produce it."""
        n = capi.lib().ccu_tape_jit_link_check(self.handle)
        if n < 0:
            raise CcuError(capi.last_error())
        return n

    def jit_chain_error(self):
        return capi.lib().ccu_tape_jit_chain_error(self.handle).decode()

    def jit_sources(self):
        L = capi.lib()
        n = L.ccu_tape_get_jit_source(self.handle, -1, None, 0)
        if n < 0:
            raise CcuError(capi.last_error())
        out = []
        for k in range(n):
            size = L.ccu_tape_get_jit_source(self.handle, k, None, 0)
            buf = ctypes.create_string_buffer(size + 1)
            L.ccu_tape_get_jit_source(self.handle, k, buf, size + 1)
            out.append(buf.value.decode())
        return out

    def program(self):
        L = capi.lib()
        n = L.ccu_tape_get_program(self.handle, None, 0)
        words = np.zeros(n, np.uint64)
        L.ccu_tape_get_program(self.handle, words.ctypes.data, n)
        return words

    def last_kernel_ms(self):
        ms = ctypes.c_double()
        capi.check(capi.lib().ccu_tape_last_kernel_ms(self.handle, ctypes.byref(ms)))
        return ms.value

    def last_eval_stats(self):
        """Phase times of the last host-pointer evaluation (ccu_tape_last_eval_stats)."""
        st = (ctypes.c_double * 6)()
        capi.check(capi.lib().ccu_tape_last_eval_stats(self.handle, st))
        return dict(h2d_ms=st[0], kernel_ms=st[1], d2h_ms=st[2], stage_ms=st[3], wall_ms=st[4], staged_bytes=st[5])

    # -- device-pointer evaluation (roofline path) --------------------------------------------------
    def eval_device(self, N, d_arg, d_res, layout=LAYOUT_SOA, stream=0, reduce_in=None, reduce_out=None):
        """d_arg/d_res: device addresses (int) or None."""
        L = capi.lib()
        a, r = capi.ptr_array(d_arg), capi.ptr_array(d_res)
        if reduce_in is None and reduce_out is None:
            capi.check(L.ccu_map_eval_device(self.handle, N, a, r, layout, stream or None))
        else:
            ri, ro = capi.int_array(reduce_in), capi.int_array(reduce_out)
            capi.check(L.ccu_map_eval_reduce_device(
                self.handle, N, a, r, None if ri is None else ri.ctypes.data_as(capi.c_i_p),
                None if ro is None else ro.ctypes.data_as(capi.c_i_p), layout, stream or None))


class CudaMap:
    """f.map(N, "cuda"): evaluates the tape for N instances; host buffers in the reference's layout
    (instance i of input j = arg[j][i*nnz_in[j] : (i+1)*nnz_in[j]], casadi/core/map.cpp:149-154)."""

    def __init__(self, tape, n, device=0, reduce_in=None, reduce_out=None, mode=None):
        if n <= 0:
            raise CcuError("Degenerate map operation")  # function.cpp:862
        self.f = tape if isinstance(tape, CudaTape) else CudaTape(tape, device, mode)
        self.n = int(n)
        self.reduce_in = list(reduce_in) if reduce_in is not None else None
        self.reduce_out = list(reduce_out) if reduce_out is not None else None

    def parallelization(self):
        return "cuda"

    def __call__(self, args, want=None):
        """args[j]: float64 array (n*nnz_in[j] values, AoS) or None.  Returns list of arrays (None where
        want[j] is False)."""
        f, N = self.f, self.n
        if len(args) != f.n_in:
            raise CcuError("expected %d inputs, got %d" % (f.n_in, len(args)))
        ins = []
        for j, a in enumerate(args):
            if a is None:
                ins.append(None)
                continue
            a = np.ascontiguousarray(a, np.float64).ravel()
            cnt = f.nnz_in[j] * (1 if (self.reduce_in and self.reduce_in[j]) else N)
            if a.size != cnt:
                raise CcuError("input %d: expected %d values, got %d" % (j, cnt, a.size))
            ins.append(a)
        outs = []
        for j in range(f.n_out):
            if want is not None and not want[j]:
                outs.append(None)
            else:
                cnt = f.nnz_out[j] * (1 if (self.reduce_out and self.reduce_out[j]) else N)
                outs.append(np.full(cnt, np.nan))
        a = capi.ptr_array([None if x is None or x.size == 0 else x.ctypes.data for x in ins])
        r = capi.ptr_array([None if x is None or x.size == 0 else x.ctypes.data for x in outs])
        L = capi.lib()
        if self.reduce_in is None and self.reduce_out is None:
            capi.check(L.ccu_map_eval_host(f.handle, N, a, r))
        else:
            ri, ro = capi.int_array(self.reduce_in), capi.int_array(self.reduce_out)
            capi.check(L.ccu_map_eval_reduce_host(
                f.handle, N, a, r, None if ri is None else ri.ctypes.data_as(capi.c_i_p),
                None if ro is None else ro.ctypes.data_as(capi.c_i_p)))
        return outs


class CudaMultiMap:
    """f.map(N, "cuda") sharded over several devices of THIS process (ccu_multi): what the C++ CudaMap builds for
    CASADI_CUDA_DEVICES.  Device g evaluates instances [g*N/G, (g+1)*N/G); reduce_out sums are merged by NCCL
    inside the library.  Host buffers in the reference's AoS layout, as for CudaMap."""

    def __init__(self, tape, n, devices, reduce_in=None, reduce_out=None, mode=None):
        L = capi.lib()
        self.nnz_in = [int(v) for v in tape["nnz_in"]]
        self.nnz_out = [int(v) for v in tape["nnz_out"]]
        self.n = int(n)
        self.reduce_in = list(reduce_in) if reduce_in is not None else None
        self.reduce_out = list(reduce_out) if reduce_out is not None else None
        arrs = [np.ascontiguousarray(tape[k], np.int32) for k in ("op", "i0", "i1", "i2")]
        d = np.ascontiguousarray(tape["d"], np.float64)
        nin, nout = np.ascontiguousarray(self.nnz_in, np.int64), np.ascontiguousarray(self.nnz_out, np.int64)
        dv = np.ascontiguousarray(list(devices), np.int32)
        capi.check(L.ccu_set_default_mode(capi.MODES[mode]))
        try:
            self.handle = L.ccu_multi_create(len(d), *[a.ctypes.data_as(capi.c_i_p) for a in arrs], d.ctypes.data_as(capi.c_d_p),
                                             int(tape["sz_w"]), len(nin), nin.ctypes.data_as(capi.c_ll_p), len(nout),
                                             nout.ctypes.data_as(capi.c_ll_p), len(dv), dv.ctypes.data_as(capi.c_i_p))
        finally:
            L.ccu_set_default_mode(-1)
        if not self.handle:
            raise CcuError(capi.last_error())

    def close(self):
        if getattr(self, "handle", None):
            capi.lib().ccu_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, args):
        N = self.n
        ins = [None if a is None else np.ascontiguousarray(a, np.float64).ravel() for a in args]
        outs = [np.full(nz * (1 if (self.reduce_out and self.reduce_out[j]) else N), np.nan) for j, nz in enumerate(self.nnz_out)]
        a = capi.ptr_array([None if x is None or x.size == 0 else x.ctypes.data for x in ins])
        r = capi.ptr_array([None if x.size == 0 else x.ctypes.data for x in outs])
        ri, ro = capi.int_array(self.reduce_in), capi.int_array(self.reduce_out)
        capi.check(capi.lib().ccu_multi_eval_host(self.handle, N, a, r, None if ri is None else ri.ctypes.data_as(capi.c_i_p),
                                                  None if ro is None else ro.ctypes.data_as(capi.c_i_p)))
        return outs
