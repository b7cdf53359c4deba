"""Random tapes through the host side of the device path, CPU only.

A generator builds random straight-line programs in the reference's tape format (ScalarAtomic stream with the
reference's habits: inputs loaded where first used, outputs stored as soon as computed, work-vector slots re-used through
a LIFO free list, sx_function.cpp:586-753), rich in what the device-side passes key on: repeated sub-expressions and
commuted operands (value numbering), constants used as operands and stored directly, inputs stored directly, unused
inputs, dead values, long dependency chains and wide independent fans (scheduler, shared-slot allocation, SPILL/FILL,
cross-segment scratch, rematerialisation).  Every tape is evaluated by the oracle on inputs that include signed zeros,
infinities and NaNs, and compared BIT for bit with
  * the interpreter program emitted by the tape compiler, executed by tests/emulator.py, for several shared-slot budgets
    and both instruction orders;
  * the CUDA source of the tape specialiser compiled with g++ (tests/test_jit_codegen.py), for several segment lengths,
    with and without rematerialisation and value numbering.
Only exactly-rounded operators are drawn, so equality of bits is the criterion.  TEST INFRASTRUCTURE."""
import os

import numpy as np
import pytest

import oracle
from casadi_b200 import CudaTape
from emulator import run_program
from test_jit_codegen import run_sources_on_host
from util import assert_bit_equal

# enum Operation (calculus.hpp:60-218): exactly-rounded subset
ADD, SUB, MUL, DIV, NEG, SQRT, SQ, TWICE = 1, 2, 3, 4, 5, 10, 11, 12
LT, LE, EQ, NE, NOT, AND, OR, FLOOR, CEIL, FABS, SIGN, COPYSIGN, IF_ELSE_ZERO, FMIN, FMAX, INV = \
    19, 20, 21, 22, 23, 24, 25, 26, 27, 29, 30, 31, 32, 34, 35, 36
CONST, INPUT, OUTPUT = 44, 45, 46
# (weights keep the values diverse: mostly ring operations, a sprinkle of the operators that collapse values to 0 / 1 / inf)
UNARY = [NEG, SQRT, SQ, TWICE, NOT, FLOOR, CEIL, FABS, SIGN, INV]
UNARY_W = np.array([4, 1.5, 0.7, 1.5, 0.2, 0.5, 0.5, 2, 0.3, 0.7])
BINARY = [ADD, SUB, MUL, DIV, LT, LE, EQ, NE, AND, OR, COPYSIGN, IF_ELSE_ZERO, FMIN, FMAX]
BINARY_W = np.array([6, 6, 3, 2, 0.3, 0.3, 0.2, 0.2, 0.15, 0.15, 1, 0.5, 1, 1])
COMMUTATIVE = {ADD, MUL}


def random_tape(rng, n_nodes, n_in, n_out_nz, shape):
    """-> tape dict.  shape: 'chain' (deep), 'fan' (wide, independent), 'mixed'."""
    nnz_in = [int(rng.integers(0, 5)) for _ in range(n_in)]
    if sum(nnz_in) == 0:
        nnz_in[0] = 3
    leaves = [("in", j, k) for j, n in enumerate(nnz_in) for k in range(n)]
    consts = [0.0, -0.0, 1.0, -1.0, 2.0, 0.5, 3.25, -7.0, 1e-3, 1e300, float(rng.normal()), float(rng.normal())]
    nodes = []  # (op, a, b) with a, b indices into `vals` (leaves first)
    vals = list(leaves) + [("const", c) for c in rng.choice(consts, size=4, replace=False)]
    n_leaf = len(vals)
    for i in range(n_nodes):
        hi = len(vals)
        def pick():
            if shape == "chain":
                return hi - 1 - int(rng.integers(0, min(3, hi))) if rng.random() < 0.8 else int(rng.integers(0, hi))
            if shape == "fan":
                return int(rng.integers(0, n_leaf)) if rng.random() < 0.5 else int(rng.integers(0, hi))
            return int(rng.integers(max(0, hi - 40), hi)) if rng.random() < 0.7 else int(rng.integers(0, hi))
        r = rng.random()
        if r < 0.12 and len(vals) > n_leaf:  # repeat an earlier operation, possibly commuted: value numbering
            op, a, b = vals[int(rng.integers(n_leaf, len(vals)))][1:]
            if op in COMMUTATIVE and rng.random() < 0.5:
                a, b = b, a
            vals.append(("op", op, a, b))
        elif r < 0.4:
            vals.append(("op", int(rng.choice(UNARY, p=UNARY_W / UNARY_W.sum())), pick(), -1))
        else:
            o, a, b = int(rng.choice(BINARY, p=BINARY_W / BINARY_W.sum())), pick(), pick()
            if o == COPYSIGN:
                # the sign bit of a NaN is a platform artefact (x86 invalid operations give -NaN, compilers fold
                # copysign(x, sqrt(y)) to fabs(x)): take the sign from b only when b is a number, else from +0
                vals.append(("op", EQ, b, b))
                vals.append(("op", IF_ELSE_ZERO, len(vals) - 1, b))
                b = len(vals) - 1
            vals.append(("op", o, a, b))
    # outputs: mostly late values, some leaves and constants stored directly, some value stored twice
    n_out = max(1, n_out_nz // 3)
    nnz_out = [0] * n_out
    stores = []
    for _ in range(n_out_nz):
        j = int(rng.integers(0, n_out))
        v = int(rng.integers(0, len(vals))) if rng.random() < 0.25 else int(rng.integers(max(n_leaf, len(vals) - 30), len(vals)))
        stores.append((v, j, nnz_out[j]))
        nnz_out[j] += 1
    return emit_tape(vals, stores, nnz_in, nnz_out)


def emit_tape(vals, stores, nnz_in, nnz_out):
    """vals: ('in', j, k) | ('const', c) | ('op', o, a, b); stores: (value, output, nonzero).  Emits the tape in the
    reference's style: depth-first from the outputs, slots from a LIFO free list, outputs stored as soon as computed."""
    uses = [0] * len(vals)
    needed = [False] * len(vals)
    stack = [v for v, _, _ in stores]
    while stack:
        v = stack.pop()
        if needed[v]:
            continue
        needed[v] = True
        if vals[v][0] == "op":
            for o in vals[v][2:]:
                if o >= 0:
                    stack.append(o)
    for v in range(len(vals)):
        if needed[v] and vals[v][0] == "op":
            for o in vals[v][2:]:
                if o >= 0:
                    uses[o] += 1
    for v, _, _ in stores:
        uses[v] += 1
    op, i0, i1, i2, d = [], [], [], [], []
    slot = {}
    free, nslots = [], 0
    by_value = {}
    for s in stores:
        by_value.setdefault(s[0], []).append(s)

    def alloc():
        nonlocal nslots
        if free:
            return free.pop()
        nslots += 1
        return nslots - 1

    def emit(o, a, b, c, x=0.0):
        op.append(o); i0.append(a); i1.append(b); i2.append(c); d.append(x)

    def release(v):
        uses[v] -= 1
        if uses[v] == 0:
            free.append(slot[v])

    def define(v):
        """post-order, iterative"""
        todo = [(v, False)]
        while todo:
            u, done = todo.pop()
            if u in slot:
                continue
            kind = vals[u][0]
            if kind == "op" and not done:
                todo.append((u, True))
                for o in reversed(vals[u][2:]):
                    if o >= 0 and o not in slot:
                        todo.append((o, False))
                continue
            if kind == "in":
                slot[u] = alloc()
                emit(INPUT, slot[u], vals[u][1], vals[u][2])
            elif kind == "const":
                slot[u] = alloc()
                emit(CONST, slot[u], 0, 0, float(vals[u][1]))
            else:
                _, o, a, b = vals[u]
                sa, sb = slot[a], slot[b] if b >= 0 else slot[a]
                release(a)
                if b >= 0:
                    release(b)
                slot[u] = alloc()  # (may be an operand's slot: the reference reads before it writes)
                emit(o, slot[u], sa, sb)
            for (_, j, k) in by_value.get(u, []):  # stored as soon as computed
                emit(OUTPUT, j, slot[u], k)
                release(u)

    for v, _, _ in stores:
        define(v)
    return {"op": np.array(op, np.int32), "i0": np.array(i0, np.int32), "i1": np.array(i1, np.int32), "i2": np.array(i2, np.int32),
            "d": np.array(d, np.float64), "sz_w": max(nslots, 1), "nnz_in": np.array(nnz_in, np.int64),
            "nnz_out": np.array(nnz_out, np.int64)}


def random_loop_tape(rng, K, ns, n_body):
    """A time-stepping tape: state0 = head(x0, p); K times state = body(state, u_k, c_k, p); outputs = tail(state_K).
    The body is one random template instantiated K times -- with an input that advances with the iteration (u_k), a
    constant that differs per iteration (c_k) and fixed operands (p, constants) -- i.e. what an expanded mapaccum /
    integrator looks like, for the re-rolling pass to find (or to refuse: either way the bits must be the reference's)."""
    nu = int(rng.integers(0, 3))
    nnz_in = [ns, 2, K * nu]  # x0, p, u (nu per iteration)
    vals = [("in", 0, i) for i in range(ns)] + [("in", 1, i) for i in range(2)]
    P = [ns, ns + 1]
    consts = {}
    def const(c):
        if c not in consts:
            vals.append(("const", c)); consts[c] = len(vals) - 1
        return consts[c]
    fixed = [const(c) for c in (0.5, 2.0, -1.25)]
    # head: mix x0 with p
    state = []
    for i in range(ns):
        vals.append(("op", ADD if i % 2 else MUL, i, P[i % 2])); state.append(len(vals) - 1)
    # body template: operands are ('s', i) state, ('u', j), ('c',) varying constant, ('p', i), ('f', i) fixed, ('b', i) body node
    ops_u = [NEG, FABS, TWICE, SQ, SQRT]
    ops_b = [ADD, SUB, MUL, DIV, FMIN, FMAX]
    tmpl = []
    def pick_operand():
        r = rng.random()
        if tmpl and r < 0.5: return ("b", int(rng.integers(max(0, len(tmpl) - 12), len(tmpl))))
        if r < 0.75: return ("s", int(rng.integers(0, ns)))
        if r < 0.82 and nu: return ("u", int(rng.integers(0, nu)))
        if r < 0.88: return ("c",)
        if r < 0.94: return ("p", int(rng.integers(0, 2)))
        return ("f", int(rng.integers(0, len(fixed))))
    for _ in range(n_body):
        if rng.random() < 0.25:
            tmpl.append((int(rng.choice(ops_u)), pick_operand(), None))
        else:
            tmpl.append((int(rng.choice(ops_b, p=[0.3, 0.3, 0.2, 0.08, 0.06, 0.06])), pick_operand(), pick_operand()))
    new_state = [int(rng.integers(max(0, n_body - 3 * ns), n_body)) for _ in range(ns)]
    h = 0.01
    for k in range(K):
        ck = const(float((k + 1) * h))  # SX folds k*h into a different constant per step
        body = []
        def ref(o):
            if o[0] == "b": return body[o[1]]
            if o[0] == "s": return state[o[1]]
            if o[0] == "u": return len_in_u(k, o[1])
            if o[0] == "c": return ck
            if o[0] == "p": return P[o[1]]
            return fixed[o[1]]
        def len_in_u(kk, j):
            key = ("in", 2, kk * nu + j)
            if key not in u_index:
                vals.append(key); u_index[key] = len(vals) - 1
            return u_index[key]
        if k == 0:
            u_index = {}
        for (o, a, b) in tmpl:
            vals.append(("op", o, ref(a), ref(b) if b is not None else -1)); body.append(len(vals) - 1)
        state = [body[i] for i in new_state]
    stores = [(state[i], 0, i) for i in range(ns)]
    vals.append(("op", ADD, state[0], state[-1])); stores.append((len(vals) - 1, 1, 0))
    return emit_tape(vals, stores, nnz_in, [ns, 1])


def random_inputs(rng, tape, N):
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-310, 1e308])
    ins = []
    for n in tape["nnz_in"]:
        a = rng.normal(size=N * int(n)) * rng.choice([1.0, 1e-3, 1e3])
        idx = rng.random(a.size) < 0.08
        a[idx] = rng.choice(special, size=int(idx.sum()))
        ins.append(a)
    return ins


CASES = [(seed, shape) for seed in range(6) for shape in ("chain", "fan", "mixed")]


def make(seed, shape):
    rng = np.random.default_rng(1000 + seed)
    tape = random_tape(rng, n_nodes=int(rng.integers(40, 700)), n_in=int(rng.integers(1, 5)),
                       n_out_nz=int(rng.integers(1, 24)), shape=shape)
    N = 33
    ins = random_inputs(rng, tape, N)
    return tape, ins, N, oracle.map_eval(tape, N, ins)


@pytest.mark.parametrize("seed,shape", CASES)
def test_random_tapes_through_the_tape_compiler(seed, shape):
    tape, ins, N, want = make(seed, shape)
    for sched in (0, 1):
        for S in (0, 3, 7, 64):
            os.environ["CCU_SCHED"] = str(sched)
            try:
                t = CudaTape(tape, device=-1)
                t.set_plan(128, 1, S)
            finally:
                os.environ.pop("CCU_SCHED", None)
            info = t.info()
            got = run_program(t.program(), N, t.nnz_in, t.nnz_out, ins, info["slots_shared"], info["slots_global"])
            for j, (g, w) in enumerate(zip(got, want)):
                assert_bit_equal(g, w, "random tape %d/%s sched=%d S=%d out%d" % (seed, shape, sched, S, j))


@pytest.mark.parametrize("seed,shape", CASES[::2])
def test_random_tapes_through_the_specialiser(seed, shape):
    tape, ins, N, want = make(seed, shape)
    for env in ({"CCU_JIT_SEG": "100000"}, {"CCU_JIT_SEG": "24"}, {"CCU_JIT_SEG": "40", "CCU_JIT_REMAT": "16"},
                {"CCU_JIT_SEG": "60", "CCU_CSE": "0", "CCU_JIT_SCHED": "0"}):
        os.environ.update(env)
        try:
            t = CudaTape(tape, device=-1)
            sources = t.jit_sources()
        finally:
            for k in env:
                os.environ.pop(k, None)
        got = run_sources_on_host(sources, t.nnz_in, t.nnz_out, ins, N)
        for j, (g, w) in enumerate(zip(got, want)):
            assert_bit_equal(g, w, "random tape %d/%s %r out%d" % (seed, shape, env, j))


@pytest.mark.parametrize("seed", range(8))
def test_random_time_stepping_tapes_flat_and_re_rolled(seed):
    """Expanded loops with advancing inputs and per-iteration constants: whatever the re-rolling pass decides (a loop
    kernel with the state in registers or in the loop scratch, or no loop), the generated code has the reference's bits."""
    rng = np.random.default_rng(7000 + seed)
    K, ns = int(rng.integers(4, 30)), int(rng.integers(1, 7))
    tape = random_loop_tape(rng, K, ns, int(rng.integers(6, 60)))
    N = 17
    ins = [rng.normal(size=N * int(n)) for n in tape["nnz_in"]]
    want = oracle.map_eval(tape, N, ins)
    found = CudaTape(tape, device=-1).loop_stats()
    for env in ({"CCU_JIT_ROLL": "0"}, {"CCU_JIT_ROLL": "1"}, {"CCU_JIT_ROLL": "1", "CCU_JIT_ROLL_REGS": "0"},
                {"CCU_JIT_ROLL": "1", "CCU_JIT_SEG": "20"}):
        os.environ.update(env)
        try:
            t = CudaTape(tape, device=-1)
            sources = t.jit_sources()
        finally:
            for k in env:
                os.environ.pop(k, None)
        got = run_sources_on_host(sources, t.nnz_in, t.nnz_out, ins, N)
        for j, (g, w) in enumerate(zip(got, want)):
            assert_bit_equal(g, w, "random loop tape %d (K=%d, loop found: %s) %r out%d" % (seed, K, found.get("found"), env, j))
