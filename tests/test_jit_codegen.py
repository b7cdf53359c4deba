"""CPU-only check of the tape specialiser's code generator (casadi_b200/csrc/jit.cpp).

The CUDA source it emits for every segment is compiled here with g++ behind a few shims (__global__,
blockIdx, __longlong_as_double ...) and executed on the host, one "thread" per instance, passing
cross-segment values through the same [slot][instance] scratch the GPU uses.  With -ffp-contract=off and
glibc's libm this reproduces the reference goldens BIT-FOR-BIT, transcendentals included, so segmentation,
scratch-slot allocation, re-materialisation and the per-operator expressions are all pinned without a GPU.
TEST INFRASTRUCTURE: nothing here is on the product path.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from casadi_b200 import CudaTape, capi, load_case, load_tape
from util import assert_bit_equal

SHIM = r"""
#include <cmath>
#include <cstring>
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
static inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
struct ccu_dim3 { unsigned x, y, z; };
static ccu_dim3 blockIdx, blockDim, threadIdx, gridDim;
// staged live-ins (TMA bulk copy -> shared memory on the device) read the scratch slot directly on the host
static double ccu_host_sm[8192];  // private shared-memory rows: one host "thread" runs at a time
#define CCU_HOST_BUILD 1
#define CCU_PF(p)
#define CCU_RING_DRAIN
#define CCU_RING_DECL
#define CCU_RING_ISSUE(row, s)
#define CCU_RING_COMMIT
#define CCU_RING_WAIT(n)
#define CCU_RING_LD(row, s) CCU_LD(s)
#define CCU_RING_ISSUE_IN(row, j, k)
#define CCU_RING_LD_IN(row, j, k) CCU_IN(j, k)
#define CCU_SM_DECL
#define CCU_SM_ST(r, v) ccu_host_sm[r] = (v)
#define CCU_SM_LD(r) ccu_host_sm[r]
#define CCU_STAGE_DECL
#define CCU_STAGE_INIT
#define CCU_STAGE_ARM(g, n)
#define CCU_STAGE_COPY(j, s, g)
#define CCU_STAGE_WAIT(g)
#define CCU_STAGE_LD(j, s) CCU_LD(s)
"""
DRIVER = r"""
extern "C" void run_seg(const ccu::IoDesc* io, long long inst0, long long n_tile, double* sc, long long flip) {
  blockDim.x = CCU_T;
  gridDim.x = (unsigned)((n_tile + CCU_T - 1) / CCU_T);
  for (long long t = 0; t < (long long)gridDim.x * CCU_T; ++t) {  // every thread of every CTA, either walking order
    blockIdx.x = (unsigned)(t / CCU_T); threadIdx.x = (unsigned)(t % CCU_T);
    ccu_seg(*io, inst0, n_tile, sc, flip);
  }
}
#ifndef CCU_TSLOTS  /* (a loop kernel names the slots of the tile scratch CCU_TSLOTS: its CCU_NSLOTS are the loop's own) */
#define CCU_TSLOTS CCU_NSLOTS
#endif
extern "C" long long scratch_doubles(long long n_tile) { return (long long)CCU_TSLOTS * CCU_T * ((n_tile + CCU_T - 1) / CCU_T); }  // CCU_SB divides CCU_T
"""


class IoDesc(ctypes.Structure):
    _fields_ = [("in_", ctypes.c_void_p * 32), ("in_si", ctypes.c_longlong * 32), ("in_sk", ctypes.c_longlong * 32),
                ("out", ctypes.c_void_p * 32), ("out_si", ctypes.c_longlong * 32), ("out_sk", ctypes.c_longlong * 32)]


def run_sources_on_host(sources, nnz_in, nnz_out, ins, N, null_in=None):
    """Compile the generated CUDA source of every segment with g++ and run it for N instances (AoS buffers)."""
    ins = [np.ascontiguousarray(a, np.float64) for a in ins]
    outs = [np.full(N * n, np.nan) for n in nnz_out]
    io = IoDesc()
    for j, a in enumerate(ins):
        io.in_[j] = None if (a.size == 0 or (null_in is not None and j in null_in)) else a.ctypes.data
        io.in_si[j], io.in_sk[j] = nnz_in[j], 1
    for j, a in enumerate(outs):
        io.out[j] = a.ctypes.data if a.size else None
        io.out_si[j], io.out_sk[j] = nnz_out[j], 1
    scratch = None
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for k, src in enumerate(sources):
            cpp = os.path.join(tmp, "seg%d.cpp" % k)
            with open(cpp, "w") as f:
                f.write(SHIM + src + DRIVER)
            procs.append(subprocess.Popen(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-w", cpp,
                                           "-o", os.path.join(tmp, "seg%d.so" % k)]))
        for p in procs:
            assert p.wait() == 0
        for k in range(len(sources)):
            lib = ctypes.CDLL(os.path.join(tmp, "seg%d.so" % k))
            if scratch is None:  # blocked [CTA][slot][thread] scratch, same size for every segment of a plan
                lib.scratch_doubles.restype = ctypes.c_longlong
                scratch = np.full(lib.scratch_doubles(ctypes.c_longlong(N)), np.nan)
            lib.run_seg(ctypes.byref(io), ctypes.c_longlong(0), ctypes.c_longlong(N),
                        scratch.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(k & 1))  # odd kernels walk in reverse
    return outs


def run_generated(tape_name, case_name, seg_instr, nmax=40, null_in=None, env=None):
    os.environ["CCU_JIT_SEG"] = str(seg_instr)
    os.environ.update(env or {})
    try:
        t = CudaTape(load_tape(tape_name), device=-1)
        sources = t.jit_sources()
    finally:
        del os.environ["CCU_JIT_SEG"]
        for k in (env or {}):
            os.environ.pop(k, None)
    case = load_case(case_name)
    N = min(case["N"], nmax)
    ins = [a[:N * n] for a, n in zip(case["in"], t.nnz_in)]
    outs = run_sources_on_host(sources, t.nnz_in, t.nnz_out, ins, N, null_in)
    return len(sources), outs, [w[:N * n] for w, n in zip(case["out"], t.nnz_out)]


@pytest.mark.parametrize("tape_name,case_name,seg", [
    ("cartpole", "cartpole", 100000), ("cartpole", "cartpole", 50), ("quad1", "quad1", 64), ("quad1", "quad1", 17),
    ("mcstep", "mcstep", 16), ("mapnode", "mapnode", 16), ("opcover", "opcover", 100000), ("opcover", "opcover", 16),
    ("opcover", "opcover_special", 20), ("quad1_jac", "quad1_jac", 300), ("mc", "mc", 500)])
def test_generated_segments_reproduce_reference_bits(tape_name, case_name, seg):
    nseg, outs, want = run_generated(tape_name, case_name, seg, nmax=700 if case_name.startswith("opcover") else 24)
    if seg < 1000:
        assert nseg > 1
    for j, (g, w) in enumerate(zip(outs, want)):
        assert_bit_equal(g, w, "%s seg=%d out%d" % (case_name, seg, j))


@pytest.mark.parametrize("tape_name,seg,remat", [("quad1_jac", 300, 24), ("quad_adj", 1200, 32), ("rocket_hess", 1200, 24),
                                                  ("opcover", 16, 12), ("mc", 500, 24)])
def test_rematerialised_plans_reproduce_reference_bits(tape_name, seg, remat):
    """Rematerialisation (tape_schedule.hpp) recomputes cross-segment values inside the reading segment: the same IEEE
    operations on the same operand values, so the outputs keep the reference's bits while the scratch traffic drops."""
    t = CudaTape(load_tape(tape_name), device=-1)
    base, re = t.jit_remat_stats(seg, 0), t.jit_remat_stats(seg, remat)
    assert re["cross_loads"] + re["cross_stores"] <= base["cross_loads"] + base["cross_stores"]
    if tape_name in ("quad_adj", "rocket_hess"):
        # reverse sweep / block Hessian: more than half of the cross-segment values are recomputed instead
        assert re["cloned"] > 0 and 2 * (re["cross_loads"] + re["cross_stores"]) < base["cross_loads"] + base["cross_stores"]
    nseg, outs, want = run_generated(tape_name, tape_name, seg, nmax=200 if tape_name == "opcover" else 12,
                                     env={"CCU_JIT_REMAT": str(remat)})
    assert nseg > 1
    for j, (g, w) in enumerate(zip(outs, want)):
        assert_bit_equal(g, w, "%s seg=%d remat=%d out%d" % (tape_name, seg, remat, j))


def test_generated_code_null_input_reads_zero():
    import oracle
    tape, case = load_tape("mapnode"), load_case("mapnode")
    N = 20
    _, outs, _ = run_generated("mapnode", "mapnode", 16, nmax=N, null_in={1})
    ins = [a[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    want = oracle.map_eval(tape, N, [ins[0], None, ins[2], ins[3]])
    for g, w in zip(outs, want):
        assert_bit_equal(g, w)


def test_jit_plan_reports_cross_segment_traffic():
    os.environ["CCU_JIT_SEG"] = "64"
    try:
        t = CudaTape(load_tape("quad1"), device=-1)
        src = t.jit_sources()
    finally:
        del os.environ["CCU_JIT_SEG"]
    assert len(src) >= 5
    assert all("ccu_seg(" in s and "--fmad" not in s for s in src)
    # every scratch slot that is read was written by an earlier segment
    written = set()
    for s in src:
        for part in s.split("CCU_LD(")[1:]:
            if part[0].isdigit():
                assert int(part.split(")")[0]) in written
        for part in s.split("CCU_ST(")[1:]:
            if part[0].isdigit():
                written.add(int(part.split(",")[0]))


def test_operand_bases_cut_the_integer_instructions(monkeypatch):
    """Per-operand base addresses (JitOptions::iobase, on by default): the kernels of an operand-heavy tape -- the
    mapaccum rollout, 400 input nonzeros -- compile (NVRTC, sm_100a, no GPU) to fewer SASS instructions than with the
    address of every access computed as i*si + k*sk, without spilling more.  tools/sass_stats.py is the same accounting;
    the timing of the A/B on a B200 is profiles/r2_sweep_iobase.jsonl."""
    import collections
    import glob
    import re
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    L = capi.lib()
    monkeypatch.setenv("CCU_JIT_CACHE", "off")

    def count(iobase):
        monkeypatch.setenv("CCU_JIT_IOBASE", str(iobase))
        t = CudaTape(load_tape("mc"), device=-1)
        ops, local = collections.Counter(), 0
        with tempfile.TemporaryDirectory() as tmp:
            n = L.ccu_tape_jit_compile_check(t.handle, tmp.encode())
            if n < 0 and "not loadable" in capi.last_error():
                pytest.skip("libnvrtc not present on this box")
            assert n > 0, capi.last_error()
            for f in glob.glob(os.path.join(tmp, "*.cubin")):
                sass = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
                for line in sass.splitlines():
                    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                    if m:
                        ops[m.group(1).split(".")[0]] += 1
        return sum(ops.values()), ops["LDL"] + ops["STL"], ops["DADD"] + ops["DMUL"] + ops["DFMA"]

    plain, plain_local, plain_fp64 = count(0)
    based, based_local, based_fp64 = count(1)
    assert based_fp64 == plain_fp64          # the arithmetic is untouched
    assert based < 0.93 * plain, (based, plain)
    assert based_local <= plain_local + 8, (based_local, plain_local)
