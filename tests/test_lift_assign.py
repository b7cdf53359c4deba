"""OP_ASSIGN (0) and OP_LIFT (88): `f = x` (casadi/core/calculus.hpp:600-606, 1006-1010).

Neither opcode can reach an SXFunction tape through the reference's own graph construction: for SXElem the templated
operators already evaluate `f = x` while the expression is built (SX::unary(OP_ASSIGN, x) and SX::binary(OP_LIFT, x, y)
return x itself), and MX `lift` expands to its first operand (mx_function.cpp:1446).  The golden `liftfun` -- made by
oracle/gen_models.cpp from exactly those constructors -- pins that.  Both opcodes are still in the dispatch switch of
SXFunction::eval (calculus.hpp:1303, 1349), so a tape handed to the C ABI may carry them: hand-made tapes are checked
against the pinned oracle (host emulation here, both kernel families under -m gpu).
"""
import os

import numpy as np
import pytest

import oracle
from casadi_b200 import CudaMap, CudaTape, load_case, load_tape
from emulator import run_program
from test_jit_codegen import run_sources_on_host
from util import assert_bit_equal

OP_ASSIGN, OP_LIFT = 0, 88


def handmade_tape():
    """y0 = lift(a*b + a, b) * 3 + lift(b, a);  y1 = lift(assign(a - b), a*b);  y2 = assign(lift(b, a)) * assign(a - b)"""
    ins = [
        (44, 0, 0, 0, 3.0),     # w0 = 3
        (45, 1, 0, 0, 0.0),     # w1 = a
        (45, 2, 1, 0, 0.0),     # w2 = b
        (3, 3, 1, 2, 0.0),      # w3 = a*b
        (1, 4, 3, 1, 0.0),      # w4 = a*b + a
        (88, 4, 4, 2, 0.0),     # w4 = lift(w4, b)
        (3, 4, 4, 0, 0.0),      # w4 = w4 * 3
        (88, 5, 2, 1, 0.0),     # w5 = lift(b, a)
        (1, 4, 4, 5, 0.0),      # w4 = w4 + w5
        (46, 0, 4, 0, 0.0),     # y0
        (2, 6, 1, 2, 0.0),      # w6 = a - b
        (0, 6, 6, 6, 0.0),      # w6 = assign(w6)
        (88, 7, 6, 3, 0.0),     # w7 = lift(w6, a*b)
        (46, 1, 7, 0, 0.0),     # y1
        (0, 5, 5, 5, 0.0),      # w5 = assign(w5)
        (3, 5, 5, 6, 0.0),      # w5 = w5 * w6
        (46, 2, 5, 0, 0.0),     # y2
    ]
    return dict(sz_w=np.int64(8), nnz_in=np.array([1, 1], np.int64), nnz_out=np.array([1, 1, 1], np.int64),
                op=np.array([r[0] for r in ins], np.int32), i0=np.array([r[1] for r in ins], np.int32),
                i1=np.array([r[2] for r in ins], np.int32), i2=np.array([r[3] for r in ins], np.int32),
                d=np.array([r[4] for r in ins], np.float64))


def inputs(N=64):
    rng = np.random.default_rng(5)
    a, b = rng.uniform(-2, 2, N), rng.uniform(-2, 2, N)
    a[3], a[7], a[20], a[40] = 0.0, -0.0, np.inf, np.nan
    b[5], b[9], b[21], b[41] = -0.0, 0.0, -np.inf, np.nan
    return [a, b]


def expected(ins):
    a, b = ins
    with np.errstate(all="ignore"):
        return [(a * b + a) * 3 + b, a - b, b * (a - b)]


def test_reference_never_emits_assign_or_lift_into_an_sx_tape():
    tape, case = load_tape("liftfun"), load_case("liftfun")
    assert OP_ASSIGN not in set(tape["op"]) and OP_LIFT not in set(tape["op"])
    outs = oracle.map_eval(tape, case["N"], case["in"])
    for j, (g, w) in enumerate(zip(outs, case["out"])):
        assert_bit_equal(g, w, "liftfun out%d" % j)
    # the golden outputs are the values of the lifted / assigned expressions themselves
    a, b = case["in"]
    with np.errstate(all="ignore"):
        assert_bit_equal(case["out"][1], a - b, "lift(assign(a-b), a*b) == a-b")


def test_oracle_evaluates_assign_and_lift_as_copies():
    tape, ins = handmade_tape(), inputs()
    outs = oracle.map_eval(tape, len(ins[0]), ins)
    for j, (g, w) in enumerate(zip(outs, expected(ins))):
        assert_bit_equal(g, w, "out%d" % j)


@pytest.mark.parametrize("S", [64, 3])
def test_compiled_program_keeps_assign_and_lift_exact(S):
    tape, ins = handmade_tape(), inputs()
    N = len(ins[0])
    t = CudaTape(tape, device=-1)
    t.set_plan(128, 1, S)
    info = t.info()
    outs = run_program(t.program(), N, t.nnz_in, t.nnz_out, ins, info["slots_shared"], info["slots_global"])
    for j, (g, w) in enumerate(zip(outs, oracle.map_eval(tape, N, ins))):
        assert_bit_equal(g, w, "S=%d out%d" % (S, j))


def test_generated_source_keeps_assign_and_lift_exact():
    tape, ins = handmade_tape(), inputs()
    N = len(ins[0])
    os.environ["CCU_JIT_SEG"] = "4"
    try:
        t = CudaTape(tape, device=-1)
        sources = t.jit_sources()
    finally:
        del os.environ["CCU_JIT_SEG"]
    outs = run_sources_on_host(sources, t.nnz_in, t.nnz_out, ins, N)
    for j, (g, w) in enumerate(zip(outs, oracle.map_eval(tape, N, ins))):
        assert_bit_equal(g, w, "generated out%d" % j)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["interp", "jit"])
def test_device_assign_and_lift_bit_exact(mode):
    tape, ins = handmade_tape(), inputs(1000)
    N = len(ins[0])
    got = CudaMap(tape, N, mode=mode)(ins)
    for j, (g, w) in enumerate(zip(got, oracle.map_eval(tape, N, ins))):
        assert_bit_equal(g, w, "%s out%d" % (mode, j))
