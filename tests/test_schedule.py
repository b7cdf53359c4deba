"""CPU-only tests of the tape scheduler (casadi_b200/csrc/tape_schedule.cpp): the min-cut bisection order is a valid
topological order (bit-exactness of the programs built from it is pinned in test_compile.py / test_jit_codegen.py),
its segments respect the requested length, and it cuts far fewer values than the reference's depth-first order
(SXFunction::init, casadi/core/sx_function.cpp:522-540) on the time-stepping tapes of BASELINE.json."""
import os

import pytest

from casadi_b200 import CudaTape, load_tape
from test_jit_codegen import run_generated
from util import assert_bit_equal


def stats(name, seg, sched, weight=0):
    """Plan statistics; `weight` = bound on a segment's estimated SASS instructions (0 = none, None = automatic)."""
    if weight is not None:
        os.environ["CCU_JIT_SEGWEIGHT"] = str(weight)
    try:
        return CudaTape(load_tape(name), device=-1).jit_plan_stats(seg, sched)
    finally:
        os.environ.pop("CCU_JIT_SEGWEIGHT", None)


@pytest.mark.parametrize("name,seg", [("quad", 800), ("quad", 2000), ("mc", 800), ("rocket_hess", 800), ("quad1_jac", 300)])
def test_bisection_cuts_fewer_values_than_the_reference_order(name, seg):
    ref, bis = stats(name, seg, 0), stats(name, seg, 1)
    assert bis["max_segment"] <= seg and ref["max_segment"] <= seg
    assert bis["cross_loads"] + bis["cross_stores"] < ref["cross_loads"] + ref["cross_stores"]
    assert bis["scratch_slots"] <= ref["scratch_slots"]


def test_quadrotor_integrator_cuts_are_the_state_vector():
    # 20 RK4 steps of a 12-state model: every cut of the bisection order carries the 12 states plus the handful of
    # step-invariant values that value numbering shares between the steps (thrust / mass, the three torques ...)
    # (reference order: 919 loads and 741 stores over 9 segments because values of all 20 steps stay alive)
    s = stats("quad", 4000, 1)
    assert s["segments"] == 2 and 12 <= s["cross_loads"] <= 18 and s["cross_stores"] == s["cross_loads"] and s["scratch_slots"] <= 18
    assert stats("quad", 8000, 1)["segments"] == 1


def test_mapaccum_tape_cuts_are_the_carried_state():
    s = stats("mc", 2000, 1)  # T=100 mapaccum of a 4-state leaf + running cost: 5 values cross the cut
    assert s["segments"] == 2 and s["cross_loads"] == 5 and s["cross_stores"] == 5


def test_schedule_is_deterministic_and_cached():
    t = CudaTape(load_tape("quad"), device=-1)
    a = t.jit_plan_stats(800, 1)
    b = t.jit_plan_stats(800, 1)
    assert {k: v for k, v in a.items() if k != "schedule_ms"} == {k: v for k, v in b.items() if k != "schedule_ms"}
    assert b["schedule_ms"] <= a["schedule_ms"]


def test_automatic_plan_sizes():
    assert stats("cartpole", 0, 1, None)["segments"] == 1
    # small tapes: kernels of <= 1200 arithmetic instructions and ~12 000 estimated SASS instructions (instruction cache):
    # the 20-step integrator with its 480 sin/cos and 640 divisions is cut into 5-9 kernels, <= 18 values per cut
    # (reference order at 9 segments: 919 + 741)
    s = stats("quad", 0, 1, None)
    assert 4 <= s["segments"] <= 12 and s["cross_loads"] <= 200 and s["cross_stores"] <= 200
    assert stats("quad", 8000, 1, 0)["segments"] == 1     # explicit plan without the code-size bound: one kernel
    # long tapes: one kernel per ~4400 instructions when the values alive inside it fit a thread's registers and spill rows
    # (the Jacobian: one RK4 step per kernel), ~2500 otherwise (reverse sweep, block Hessian)
    j = stats("quad_jac", 0, 1, None)
    assert j["segments"] <= 24 and j["max_live"] <= 170 and j["cross_loads"] + j["cross_stores"] < 4500
    a = stats("quad_adj", 0, 1, None)
    assert a["max_segment"] <= 2500 and a["segments"] >= 8
    assert stats("mc", 0, 1, None)["segments"] >= 2
