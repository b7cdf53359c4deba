"""Loop re-rolling (casadi_b200/csrc/tape_reroll.cpp) and the re-rolled plans of the specialising code generator.

An expanded mapaccum / fold tower (function.cpp:692-746) or an SX integrator and its AD products are T copies of one
step on the tape.  The pass recovers the loop from the value graph, verifies operand by operand that iterating the body
reproduces the tape, and jit.cpp then emits ONE persistent loop kernel (state in registers or in an L2-resident per-CTA
scratch) between the kernels of the irregular head and tail.  Everything here runs without a GPU: the generated CUDA
source is compiled with g++ and executed on the host (test_jit_codegen.py), where it must reproduce the reference
goldens bit for bit; the -m gpu tests repeat that on the device against the flat plan.
"""
import os

import numpy as np
import pytest

from casadi_b200 import CudaMap, CudaTape, capi, load_case, load_tape
from test_jit_codegen import run_generated
from util import assert_bit_equal


def loop(name):
    return CudaTape(load_tape(name), device=-1).loop_stats()


def test_loops_of_the_baseline_tapes():
    q = loop("quad")       # 20 RK4 steps of the 12-state quadrotor: the first step is folded with the inputs
    assert q["found"] and q["iters"] == 19 and q["body"] == 284 and q["carried"] == 12 and q["exits"] == 12
    j = loop("quad_jac")   # its Jacobian: 18 identical sensitivity steps; SX folded the constant blocks (k * h), which
    assert j["found"] and j["iters"] == 18 and j["body"] > 3000 and j["varying_constants"] > 0   # differ per iteration
    assert j["iters"] * j["body"] > 0.85 * 67811
    f = loop("quad_fwd")
    assert f["found"] and f["iters"] >= 17 and f["carried"] >= 24
    m = loop("mc")         # mapaccum T=100: the noise inputs advance with the iteration
    assert m["found"] and m["iters"] == 99 and m["carried"] == 5 and m["advancing_inputs"] > 0


@pytest.mark.parametrize("name", ["opcover", "quad1", "quad1_jac", "rocket_hess", "quad_adj", "mcstep", "mapnode"])
def test_tapes_without_a_forward_loop_stay_flat(name):
    st = loop(name)
    assert not st["found"] and st["why"].startswith("no loop")
    os.environ["CCU_JIT_ROLL"] = "1"
    try:
        src = CudaTape(load_tape(name), device=-1).jit_sources()
    finally:
        del os.environ["CCU_JIT_ROLL"]
    assert all("CCU_KERNEL_LOOP" not in s for s in src)


@pytest.mark.parametrize("name,seg,env", [
    ("quad", 1200, {}),                                # state in registers
    ("quad", 1200, {"CCU_JIT_ROLL_REGS": "0"}),        # state double-buffered in the loop scratch
    ("quad", 100, {}),                                 # body cut into several segments
    ("mc", 1200, {}),                                  # advancing inputs
    ("mc", 1200, {"CCU_JIT_ROLL_REGS": "0"}),
    ("quad_fwd", 300, {}),
    ("quad_jac", 4400, {}),                            # one segment per step, 32 per-iteration constants
    ("quad_jac", 1200, {"CCU_JIT_FASTOPS": "0"}),
])
def test_rolled_plans_reproduce_reference_bits(name, seg, env):
    e = dict(env, CCU_JIT_ROLL="1")
    nseg, outs, want = run_generated(name, name, seg, nmax=6 if name == "quad_jac" else 24, env=e)
    os.environ.update(e)
    os.environ["CCU_JIT_SEG"] = str(seg)
    try:
        src = CudaTape(load_tape(name), device=-1).jit_sources()
    finally:
        for k in list(e) + ["CCU_JIT_SEG"]:
            os.environ.pop(k, None)
    assert sum("CCU_KERNEL_LOOP" in s for s in src) == 1 and len(src) == nseg
    for j, (g, w) in enumerate(zip(outs, want)):
        assert_bit_equal(g, w, "%s rolled seg=%d out%d" % (name, seg, j))


def test_rolled_kernels_compile_for_sm_100a():
    """NVRTC needs no GPU: the loop kernel (register variant and scratch variant) builds for the device."""
    L = capi.lib()
    for env in ({}, {"CCU_JIT_ROLL_REGS": "0"}):
        os.environ.update(dict(env, CCU_JIT_ROLL="1"))
        try:
            t = CudaTape(load_tape("quad"), device=-1)
            n = L.ccu_tape_jit_compile_check(t.handle, None)
        finally:
            for k in list(env) + ["CCU_JIT_ROLL"]:
                os.environ.pop(k, None)
        if n < 0 and "not loadable" in capi.last_error():
            pytest.skip("libnvrtc not present on this box")
        assert n > 0, capi.last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("name,env", [("quad", {}), ("quad", {"CCU_JIT_ROLL_REGS": "0"}), ("mc", {}), ("quad_fwd", {}),
                                      ("quad_jac", {})])
def test_device_rolled_plan_has_the_bits_of_the_flat_plan(name, env):
    tape, case = load_tape(name), load_case(name)
    P = case["N"]
    reps = 700 if name != "quad_jac" else 1200          # several blocks per resident CTA, a ragged last block
    N = P * reps - 3
    ins = [np.tile(a, reps)[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    os.environ["CCU_JIT_ROLL"] = "0"
    try:
        flat = CudaMap(tape, N, mode="jit")
    finally:
        del os.environ["CCU_JIT_ROLL"]
    want = flat(ins)
    os.environ.update(dict(env, CCU_JIT_ROLL="1"))
    try:
        rolled = CudaMap(tape, N, mode="jit")
    finally:
        for k in list(env) + ["CCU_JIT_ROLL"]:
            os.environ.pop(k, None)
    assert rolled.f.info()["jit_loop_iters"] > 0 and flat.f.info()["jit_loop_iters"] == 0
    got = rolled(ins)
    for j, (g, w) in enumerate(zip(got, want)):
        assert_bit_equal(g, w, "%s rolled out%d" % (name, j))
    for j, (g, w) in enumerate(zip(got, case["out"])):
        err = np.abs(g[:len(w)] - w) / np.maximum(np.abs(w), 1.0)
        assert np.nanmax(err, initial=0.0) <= 1e-11
