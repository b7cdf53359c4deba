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
    # every kernel stays within ~7000 estimated SASS instructions (instruction cache): the 20-step integrator with its
    # 480 sin/cos and 640 divisions is cut into 9 kernels, ~18 values per cut (reference order at 9 segments: 919 + 741)
    s = stats("quad", 0, 1, None)
    assert 4 <= s["segments"] <= 12 and s["cross_loads"] <= 200 and s["cross_stores"] <= 200
    assert stats("quad", 0, 1, 0)["segments"] == 1        # without that bound: <= 8000 arithmetic instructions, one kernel
    s = stats("rocket_hess", 0, 1, None)                  # longer: cut every <= 2500
    assert s["segments"] > 1 and s["max_segment"] <= 2500


@pytest.mark.parametrize("sched", ["0", "1"])
@pytest.mark.parametrize("opts", [{}, {"CCU_JIT_REGVALS": "6", "CCU_JIT_SPILL": "-1"}, {"CCU_JIT_REGVALS": "5", "CCU_JIT_SPILL": "3", "CCU_JIT_STAGE": "2"},
                                  {"CCU_JIT_STAGE": "-1"}])
def test_generated_code_variants_reproduce_reference_bits(sched, opts):
    """Reference / bisection order x (plain, shared-memory spill rows, spill rows + staged live-ins + compiler-managed
    overflow, everything staged): the generated segments reproduce the reference bits on the host."""
    env = dict(opts, CCU_JIT_SCHED=sched)
    os.environ.update(env)
    try:
        for tape, seg in (("quad1_jac", 300), ("mc", 500)):
            nseg, outs, want = run_generated(tape, tape, seg, nmax=12)
            assert nseg > 1
            for j, (g, w) in enumerate(zip(outs, want)):
                assert_bit_equal(g, w, "%s %s out%d" % (tape, env, j))
    finally:
        for k in env:
            os.environ.pop(k, None)


def test_chain_kernel_links_on_the_host():
    """The persistent chain kernel (every segment a relocatable device function, linked by nvJitLink for sm_100a)
    builds without a GPU; skipped when NVRTC / nvJitLink are not loadable."""
    os.environ["CCU_JIT_SEG"] = "300"
    try:
        t = CudaTape(load_tape("quad1_jac"), device=-1)
        try:
            n = t.jit_link_check()
        except Exception as e:  # noqa: BLE001
            if "not loadable" in str(e):
                pytest.skip(str(e))
            raise
    finally:
        del os.environ["CCU_JIT_SEG"]
    assert n > 10000


def test_parallel_recursion_is_deterministic():
    """The bisection recursion runs its halves (and the two cuts of a large piece) on several threads; the order it
    produces must not depend on how many."""
    res = []
    for th in ("1", "4", "16"):
        os.environ["CCU_SCHED_THREADS"] = th
        try:
            t = CudaTape(load_tape("quad_fwd"), device=-1)
            res.append(["".join(t.jit_sources())])
        finally:
            del os.environ["CCU_SCHED_THREADS"]
    assert res[0] == res[1] == res[2]
