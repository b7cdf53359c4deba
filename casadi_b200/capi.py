"""ctypes binding of libcasadi_cuda.so (include/casadi_cuda.h).

The product path: everything here calls the CUDA library through its C ABI.  There is no CPU
fallback -- if the shared library is missing, or no CUDA device is present, the calls raise.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcasadi_cuda.so")

c_ll = ctypes.c_longlong
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
c_i_p = ctypes.POINTER(ctypes.c_int)
c_d_p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

LAYOUT_AOS = 0
LAYOUT_SOA = 1
MODE_INTERP = 0
MODE_JIT = 1
MODE_AUTO = 2
MODES = {"interp": MODE_INTERP, "jit": MODE_JIT, "auto": MODE_AUTO, None: -1}

# every symbol include/casadi_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ccu_abi_version": (ctypes.c_int, []),
    "ccu_last_error": (ctypes.c_char_p, []),
    "ccu_device_count": (ctypes.c_int, []),
    "ccu_tape_create": (c_vp, [c_ll, c_i_p, c_i_p, c_i_p, c_i_p, c_d_p, c_ll, c_ll, c_ll_p, c_ll, c_ll_p,
                               ctypes.c_int]),
    "ccu_tape_destroy": (None, [c_vp]),
    "ccu_tape_get_info": (ctypes.c_int, [c_vp, c_vp]),
    "ccu_tape_get_program": (c_ll, [c_vp, c_vp, c_ll]),
    "ccu_tape_set_plan": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "ccu_tape_set_mode": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "ccu_set_default_mode": (ctypes.c_int, [ctypes.c_int]),
    "ccu_tape_set_jit_plan": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ll]),
    "ccu_tape_get_jit_source": (c_ll, [c_vp, c_ll, c_vp, c_ll]),
    "ccu_tape_jit_plan_stats": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_ll_p]),
    "ccu_tape_jit_remat_stats": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_ll_p]),
    "ccu_tape_set_jit_remat": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "ccu_tape_loop_stats": (ctypes.c_int, [c_vp, c_ll_p]),
    "ccu_tape_set_jit_schedule": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "ccu_tape_jit_link_check": (c_ll, [c_vp]),
    "ccu_tape_jit_chain_error": (ctypes.c_char_p, [c_vp]),
    "ccu_tape_jit_compile_check": (c_ll, [c_vp, ctypes.c_char_p]),
    "ccu_map_eval_host": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp]),
    "ccu_map_eval_device": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, ctypes.c_int, c_vp]),
    "ccu_map_eval_reduce_host": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, c_i_p, c_i_p]),
    "ccu_map_eval_reduce_device": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, c_i_p, c_i_p, ctypes.c_int, c_vp]),
    "ccu_map_eval_shard_device": (ctypes.c_int, [c_vp, c_ll, c_ll, c_ll, c_vp, c_vp, c_i_p, c_i_p, c_vp, ctypes.c_int, c_vp]),
    "ccu_reduce_tree_device": (ctypes.c_int, [ctypes.c_int, c_vp, c_ll, c_ll, c_vp, c_vp]),
    "ccu_comm_available": (ctypes.c_int, []),
    "ccu_comm_nccl_version": (ctypes.c_int, []),
    "ccu_comm_create_all": (c_vp, [ctypes.c_int, c_i_p]),
    "ccu_comm_unique_id": (ctypes.c_int, [c_vp]),
    "ccu_comm_create_rank": (c_vp, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "ccu_comm_destroy": (None, [c_vp]),
    "ccu_comm_size": (ctypes.c_int, [c_vp]),
    "ccu_comm_allreduce_block_sums": (ctypes.c_int, [c_vp, c_vp, c_ll, c_vp]),
    "ccu_multi_create": (c_vp, [c_ll, c_i_p, c_i_p, c_i_p, c_i_p, c_d_p, c_ll, c_ll, c_ll_p, c_ll, c_ll_p, ctypes.c_int, c_i_p]),
    "ccu_builder_finish_multi": (c_vp, [c_vp, c_ll, c_ll_p, c_ll, c_ll_p, ctypes.c_int, c_i_p]),
    "ccu_multi_destroy": (None, [c_vp]),
    "ccu_multi_size": (ctypes.c_int, [c_vp]),
    "ccu_multi_tape": (c_vp, [c_vp, ctypes.c_int]),
    "ccu_multi_eval_host": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, c_i_p, c_i_p]),
    "ccu_multi_eval_host_grouped": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, c_i_p, c_i_p, c_i_p, c_i_p]),
    "ccu_builder_create": (c_vp, []),
    "ccu_builder_destroy": (None, [c_vp]),
    "ccu_builder_const": (c_ll, [c_vp, ctypes.c_double]),
    "ccu_builder_input": (c_ll, [c_vp, c_ll, c_ll]),
    "ccu_builder_op": (c_ll, [c_vp, ctypes.c_int, c_ll, c_ll]),
    "ccu_builder_output": (ctypes.c_int, [c_vp, c_ll, c_ll, c_ll]),
    "ccu_builder_ldl": (ctypes.c_int, [c_vp, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll, c_ll_p]),
    "ccu_builder_qr": (ctypes.c_int, [c_vp, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll, ctypes.c_int,
                                      ctypes.c_double, c_ll_p]),
    "ccu_builder_mtimes": (ctypes.c_int, [c_vp, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p]),
    "ccu_builder_select": (c_ll, [c_vp, c_ll, c_ll, c_ll]),
    "ccu_builder_export": (c_ll, [c_vp, c_i_p, c_i_p, c_i_p, c_i_p, c_d_p, c_ll, c_ll_p]),
    "ccu_builder_finish": (c_vp, [c_vp, c_ll, c_ll_p, c_ll, c_ll_p, ctypes.c_int]),
    "ccu_ldl_create": (c_vp, [c_ll_p, c_ll_p, c_ll_p, c_ll, ctypes.c_int]),
    "ccu_qr_create": (c_vp, [c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll_p, c_ll, ctypes.c_int, ctypes.c_double, ctypes.c_int]),
    "ccu_linsol_destroy": (None, [c_vp]),
    "ccu_linsol_tape": (c_vp, [c_vp]),
    "ccu_linsol_solve_host": (ctypes.c_int, [c_vp, c_ll, c_d_p, c_d_p, c_d_p, c_ll_p]),
    "ccu_linsol_solve_device": (ctypes.c_int, [c_vp, c_ll, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp]),
    "ccu_tape_last_kernel_ms": (ctypes.c_int, [c_vp, c_d_p]),
    "ccu_tape_last_eval_stats": (ctypes.c_int, [c_vp, c_d_p]),
    "ccu_launch_count": (c_ll, []),
    "ccu_fp64_issue_rate": (ctypes.c_int, [ctypes.c_int, c_d_p]),
    "ccu_selftest_fastops": (ctypes.c_int, [ctypes.c_int, c_ll, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)]),
    "ccu_selftest_host_copy": (ctypes.c_int, [c_ll, ctypes.c_int, ctypes.c_int, c_d_p]),
    "ccu_set_device": (ctypes.c_int, [ctypes.c_int]),
    "ccu_malloc": (c_vp, [c_ll]),
    "ccu_free": (ctypes.c_int, [c_vp]),
    "ccu_malloc_host": (c_vp, [c_ll]),
    "ccu_free_host": (ctypes.c_int, [c_vp]),
    "ccu_host_register": (ctypes.c_int, [c_vp, c_ll]),
    "ccu_host_unregister": (ctypes.c_int, [c_vp]),
    "ccu_host_registered_count": (ctypes.c_int, []),
    "ccu_memcpy_h2d": (ctypes.c_int, [c_vp, c_vp, c_ll, c_vp]),
    "ccu_memcpy_d2h": (ctypes.c_int, [c_vp, c_vp, c_ll, c_vp]),
    "ccu_stream_sync": (ctypes.c_int, [c_vp]),
    "ccu_device_sync": (ctypes.c_int, []),
}


class TapeInfo(ctypes.Structure):
    _fields_ = [(n, c_ll) for n in ("n_instr", "n_words", "flops", "bytes_in", "bytes_out", "sz_w", "slots_shared",
                                    "slots_global", "threads", "ipt", "smem_bytes", "spill_loads", "spill_stores",
                                    "max_live", "grid", "ctas_per_sm", "mode", "jit_segments",
                                    "jit_scratch_slots", "jit_tile", "jit_compile_ms", "jit_cross_loads",
                                    "jit_cross_stores", "jit_max_regs", "jit_cache_hits", "jit_threads",
                                    "jit_schedule", "jit_schedule_ms", "jit_chained", "cse_removed", "jit_remat_cloned", "jit_loop_iters",
                                    "jit_loop_body", "jit_loop_slots")]


_lib = None


class CcuError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CcuError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().ccu_last_error().decode()


def check(rc):
    if rc != 0:
        raise CcuError(last_error())


def ptr_array(ptrs):
    """(void*)[n] from ints/None."""
    n = max(len(ptrs), 1)
    return (ctypes.c_void_p * n)(*[None if (p is None or p == 0) else int(p) for p in ptrs])


def int_array(vals):
    if vals is None:
        return None
    a = np.ascontiguousarray([1 if v else 0 for v in vals], np.int32)
    return a


def selftest_host_copy(nbytes, dst_misalign=0, src_misalign=0):
    """GB/s of the staging copy of the host path (raises when the copy is not exact)."""
    r = ctypes.c_double(0)
    check(lib().ccu_selftest_host_copy(nbytes, dst_misalign, src_misalign, ctypes.byref(r)))
    return r.value


def selftest_fastops(n, seed=1, device=0):
    """(mismatches, flagged, checks) of ccu_selftest_fastops."""
    c = (ctypes.c_ulonglong * 3)()
    check(lib().ccu_selftest_fastops(device, n, seed, c))
    return int(c[0]), int(c[1]), int(c[2])
