"""CPU-only tests of the host side: C-ABI library loads and exports every declared symbol, the tape
compiler (SSA recovery, shared-slot allocation, SPILL/FILL, register forwarding) preserves the
reference results bit-for-bit (checked by emulating the emitted program), error behaviour."""
import os
import re

import numpy as np
import pytest

from casadi_b200 import CcuError, CudaMap, CudaTape, capi, load_case, load_tape
from emulator import run_program
from util import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "casadi_cuda.h")).read()
    declared = set(re.findall(r"CCU_EXPORT[^;(]*?\b(ccu_\w+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libcasadi_cuda.so does not export %s" % name
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert L.ccu_abi_version() == 3


def emulate(tape_name, case_name, S, nmax=64, sched=None):
    tape = load_tape(tape_name)
    case = load_case(case_name)
    N = min(case["N"], nmax)
    if sched is not None:
        os.environ["CCU_SCHED"] = str(sched)
    try:
        t = CudaTape(tape, device=-1)
        t.set_plan(128, 1, S)
    finally:
        os.environ.pop("CCU_SCHED", None)
    info = t.info()
    ins = [a[:N * n] for a, n in zip(case["in"], t.nnz_in)]
    outs = run_program(t.program(), N, t.nnz_in, t.nnz_out, ins, info["slots_shared"], info["slots_global"])
    for j, (got, want) in enumerate(zip(outs, case["out"])):
        assert_bit_equal(got, want[:N * t.nnz_out[j]], "%s S=%d out%d" % (case_name, S, j))
    return info


@pytest.mark.parametrize("name", ["cartpole", "cartpole1", "quad1", "mcstep", "mapnode", "opcover"])
def test_compiled_program_small_tapes_no_spill(name):
    info = emulate(name, name, 64)  # every work vector here has at most 58 live values
    assert info["slots_global"] == 0


def test_automatic_plan_small_window_for_large_work_vectors():
    # the automatic plan keeps <= 32 live values entirely in shared memory and otherwise uses a 16-slot window
    assert emulate("cartpole", "cartpole", 0)["slots_global"] == 0
    info = emulate("quad1", "quad1", 0)
    assert info["slots_shared"] == 16 and info["slots_global"] > 0


@pytest.mark.parametrize("sched", [0, 1])
@pytest.mark.parametrize("name,S", [("cartpole", 4), ("cartpole", 8), ("cartpole", 16), ("quad1", 6), ("quad1", 24),
                                    ("quad", 48), ("quad", 16), ("mc", 8), ("quad1_jac", 32), ("mapnode", 4)])
def test_compiled_program_with_spills(name, S, sched):
    """Both instruction orders (reference depth-first / min-cut bisection) reproduce the reference bits."""
    info = emulate(name, name, S, nmax=24, sched=sched)
    if name != "mapnode":
        assert info["slots_global"] > 0 and info["spill_loads"] > 0


def test_bisection_order_shrinks_the_live_set():
    # mapaccum T=100 (reference order: 388 values spilled with a 16-slot window) fits 32 shared slots entirely, and
    # the 20-step quadrotor integrator (591 live values in reference order) fits 64
    assert emulate("mc", "mc", 32, nmax=8, sched=1)["slots_global"] == 0
    assert emulate("mc", "mc", 32, nmax=8, sched=0)["slots_global"] > 0
    assert emulate("quad", "quad", 64, nmax=8, sched=1)["slots_global"] == 0


@pytest.mark.parametrize("name", ["quad_fwd", "quad_adj", "rocket_hess"])
def test_compiled_program_big_tapes(name):
    emulate(name, name, 48, nmax=8)


def test_compiled_program_special_values():
    emulate("opcover", "opcover_special", 4, nmax=700)


def test_no_acc_forwarding_gives_same_results(monkeypatch):
    monkeypatch.setenv("CCU_NO_ACC", "1")
    emulate("quad1", "quad1", 12, nmax=16)


def _mk(op, i0, i1, i2, d, sz_w, nnz_in, nnz_out):
    return dict(op=np.array(op, np.int32), i0=np.array(i0, np.int32), i1=np.array(i1, np.int32),
                i2=np.array(i2, np.int32), d=np.array(d, np.float64), sz_w=sz_w,
                nnz_in=np.array(nnz_in, np.int64), nnz_out=np.array(nnz_out, np.int64))


@pytest.mark.parametrize("bad_op,msg", [(48, "OP_CALL"), (47, "OP_PARAMETER"), (87, "OP_PRINTME"), (52, "not a scalar")])
def test_unsupported_opcodes_fail_loudly_at_create(bad_op, msg):
    tape = _mk([45, bad_op, 46], [0, 1, 0], [0, 0, 1], [0, 0, 0], [0, 0, 0], 2, [1], [1])
    with pytest.raises(CcuError, match=msg):
        CudaTape(tape, device=-1)


def test_out_of_range_indices_rejected():
    with pytest.raises(CcuError, match="out of range"):
        CudaTape(_mk([45, 46], [0, 0], [0, 0], [3, 0], [0, 0], 1, [1], [1]), device=-1)
    with pytest.raises(CcuError, match="out of range"):
        CudaTape(_mk([45, 1, 46], [0, 5, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0], 2, [1], [1]), device=-1)
    with pytest.raises(CcuError, match="undefined slot"):
        CudaTape(_mk([45, 1, 46], [0, 1, 0], [0, 0, 1], [0, 1, 0], [0, 0, 0], 2, [1], [1]), device=-1)


def test_eval_without_device_fails_loudly():
    t = CudaTape(load_tape("cartpole1"), device=-1)
    m = CudaMap(t, 4)
    with pytest.raises(CcuError, match="no CPU fallback"):
        m([np.zeros(16), np.zeros(4)])


def test_no_gpu_means_create_fails_not_falls_back():
    if capi.lib().ccu_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(CcuError, match="no CPU fallback"):
        CudaTape(load_tape("cartpole1"), device=0)


def test_degenerate_map_rejected():
    t = CudaTape(load_tape("cartpole1"), device=-1)
    with pytest.raises(CcuError, match="Degenerate"):
        CudaMap(t, 0)


def test_empty_tape_and_zero_nnz_io():
    t = CudaTape(_mk([], [], [], [], [], 0, [0, 2], [0]), device=-1)
    assert t.info()["n_words"] == 1


@pytest.mark.parametrize("nbytes,dmis,smis", [(0, 0, 0), (1, 0, 0), (4095, 3, 5), (300 << 10, 8, 0), (300 << 10, 24, 16),
                                               ((5 << 20) + 1237, 8, 8), ((9 << 20) + 8, 0, 40), (64 << 20, 16, 0)])
def test_host_staging_copy_is_exact(nbytes, dmis, smis):
    """The multi-threaded streaming copy of the host path (csrc/hostcopy.cpp, CopyPool) moves exactly the bytes it is
    given, whatever the alignment of the caller's buffer (a std::vector<double> is 16-byte aligned, a slice of one 8)."""
    rate = capi.selftest_host_copy(nbytes, dmis, smis)
    assert rate >= 0
