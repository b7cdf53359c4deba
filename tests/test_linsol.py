"""Batched LDL / QR with shared sparsity (ccu_ldl_create / ccu_qr_create): the factorisation and the solves are
TRACED over the shared pattern into an ordinary tape (casadi_b200/csrc/tape_builder.cpp), so the tests are
  * CPU: the traced tape, run through the generated-code host harness and through the device-program emulator,
    against the reference's own `solve(K,b,"ldl"|"qr")` outputs (tests/golden/kkt_*.case.npz) -- bit-exact;
  * GPU: ccu_linsol_solve_host in both execution modes against the same goldens -- bit-exact."""
import os

import numpy as np
import pytest

import oracle
from casadi_b200 import CudaLinsol, capi, load_case
from casadi_b200.tapeio import GOLDEN_DIR
from emulator import run_program
from test_jit_codegen import run_sources_on_host
from util import assert_bit_equal


def kkt_sym():
    z = np.load(os.path.join(GOLDEN_DIR, "kkt.sym.npz"))
    return {k: np.array(z[k]) for k in z.files}


def make(kind, device, mode=None, nrhs=1):
    s = kkt_sym()
    sym = (s["sp_lt"], s["p"]) if kind == "ldl" else (s["sp_v"], s["sp_r"], s["prinv"], s["pc"])
    return CudaLinsol(kind, s["sp_a"], sym, nrhs=nrhs, device=device, mode=mode)


@pytest.mark.parametrize("kind", ["ldl", "qr"])
def test_traced_tape_generated_code_matches_reference_solve(kind, monkeypatch):
    monkeypatch.setenv("CCU_JIT_SEG", "700")
    ls = make(kind, -1)
    info = ls.tape.info()
    assert info["flops"] > 1000
    case = load_case("kkt_" + kind)
    N = 40
    A, B = case["in"][0][:N * ls.nnz_a], case["in"][1][:N * 60]
    src = ls.tape.jit_sources()
    assert len(src) > 1
    outs = run_sources_on_host(src, ls.tape.nnz_in, ls.tape.nnz_out, [A, B], N)
    assert_bit_equal(outs[0], case["out"][0][:N * 60], "%s x" % kind)
    assert not outs[1].any()  # no zero pivot / no singular R in these systems


@pytest.mark.parametrize("kind", ["ldl", "qr"])
def test_traced_tape_device_program_emulated(kind):
    ls = make(kind, -1)
    ls.tape.set_plan(128, 1, 40)  # force SPILL/FILL traffic through the allocator
    info = ls.tape.info()
    case = load_case("kkt_" + kind)
    N = 6
    A, B = case["in"][0][:N * ls.nnz_a], case["in"][1][:N * 60]
    outs = run_program(ls.tape.program(), N, ls.tape.nnz_in, ls.tape.nnz_out, [A, B], info["slots_shared"], info["slots_global"])
    assert_bit_equal(outs[0], case["out"][0][:N * 60], "%s x (emulated device program)" % kind)


def test_singular_systems_are_flagged_like_the_reference():
    """LinsolQr::nfact fails when |R_cc| < eps (linsol_qr.cpp:146-163); LinsolLdl::nfact warns on zeros in D
    (linsol_ldl.cpp:122-124).  The traced tapes count such instances; checked against the oracle's casadi_qr_singular."""
    s = kkt_sym()
    case = load_case("kkt_qr")
    nnz = int(s["sp_a"][2 + 60])
    A = case["in"][0][:3 * nnz].reshape(3, nnz).copy()
    A[1] = 0.0  # a zero matrix is singular whatever the permutation
    _, _, r, _ = oracle.qr_factor_solve(s["sp_a"], s["sp_v"], s["sp_r"], s["prinv"], s["pc"], A[1], np.zeros(60))
    assert oracle.lib().oracle_qr_singular(None, None, r.ctypes.data_as(oracle.c_d_p), s["sp_r"].ctypes.data_as(oracle.c_ll_p),
                                           s["pc"].ctypes.data_as(oracle.c_ll_p), __import__("ctypes").c_double(1e-12)) > 0
    for kind in ("qr", "ldl"):
        ls = make(kind, -1)
        src = ls.tape.jit_sources()
        outs = run_sources_on_host(src, ls.tape.nnz_in, ls.tape.nnz_out, [A.ravel(), case["in"][1][:180]], 3)
        assert list(outs[1]) == [0.0, 1.0, 0.0], (kind, outs[1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["interp", "jit"])
@pytest.mark.parametrize("kind", ["ldl", "qr"])
def test_gpu_batched_solve_bit_exact_vs_reference(kind, mode):
    ls = make(kind, 0, mode)
    assert ls.tape.info()["mode"] == capi.MODES[mode]
    case = load_case("kkt_" + kind)
    X, flagged = ls.solve(case["in"][0], case["in"][1])
    assert flagged == 0
    assert_bit_equal(X.ravel(), case["out"][0], "%s/%s x" % (kind, mode))


@pytest.mark.gpu
def test_gpu_batched_solve_large_batch_periodic_and_flags():
    """1e5 systems (periodic inputs => periodic outputs, first period = reference golden), a few made singular."""
    ls = make("qr", 0)
    case = load_case("kkt_qr")
    P, reps = case["N"], 700
    A = np.tile(case["in"][0].reshape(P, -1), (reps, 1))
    B = np.tile(case["in"][1].reshape(P, -1), (reps, 1))
    A[[5, 77777, 104999]] = 0.0
    X, flagged = ls.solve(A, B)
    assert flagged == 3
    X = X.reshape(reps, P, 60)
    good = np.ones((reps, P), bool).ravel()
    good[[5, 77777, 104999]] = False
    ref = np.broadcast_to(case["out"][0].reshape(P, 60), (reps, P, 60)).reshape(-1, 60)
    assert_bit_equal(X.reshape(-1, 60)[good], ref[good], "periodic batch")
