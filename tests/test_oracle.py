"""Pin the CPU oracle (oracle/oracle.c) bit-for-bit against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by oracle/gen_models.cpp linked against the
reference itself: outputs of f.map(N,"serial") (casadi/core/map.cpp:141-157) on seeded inputs.
"""
import numpy as np
import pytest

import oracle
from casadi_b200.tapeio import load_case, load_tape
from util import assert_bit_equal

SX_CASES = [("cartpole", "cartpole"), ("cartpole1", "cartpole1"), ("quad", "quad"), ("quad1", "quad1"),
            ("quad_fwd", "quad_fwd"), ("quad_adj", "quad_adj"), ("quad_jac", "quad_jac"),
            ("quad1_jac", "quad1_jac"), ("rocket_hess", "rocket_hess"), ("mcstep", "mcstep"), ("mc", "mc"),
            ("mapnode", "mapnode"), ("opcover", "opcover"), ("opcover", "opcover_special")]


@pytest.mark.parametrize("tape_name,case_name", SX_CASES)
def test_oracle_matches_reference_serial_map(tape_name, case_name):
    tape = load_tape(tape_name)
    case = load_case(case_name)
    outs = oracle.map_eval(tape, case["N"], case["in"])
    for j, (got, want) in enumerate(zip(outs, case["out"])):
        assert_bit_equal(got, want, "%s out%d" % (case_name, j))


def test_oracle_null_input_reads_zero_and_null_output_skipped():
    tape = load_tape("mapnode")
    case = load_case("mapnode")
    N = case["N"]
    ins = list(case["in"])
    zero = [np.zeros_like(a) for a in ins]
    full = oracle.map_eval(tape, N, [ins[0], zero[1], ins[2], ins[3]])
    part = oracle.map_eval(tape, N, [ins[0], None, ins[2], ins[3]], want=[True, False, True])
    assert part[1] is None
    assert_bit_equal(part[0], full[0])
    assert_bit_equal(part[2], full[2])


def test_oracle_repsum_matches_reference_mapsum():
    tape = load_tape("mc")
    case = load_case("mc")
    ref = load_case("mc_sum")
    outs = oracle.map_eval(tape, case["N"], case["in"])
    assert_bit_equal(oracle.repsum(outs[0], 4, case["N"]), ref["out"][0], "sum xT")
    assert_bit_equal(oracle.repsum(outs[1], 1, case["N"]), ref["out"][1], "sum J")


def test_ulp_distance_helper_resolves_single_ulps():
    from util import ulp_diff
    a = np.array([1.0, 1.1935021221110558, -3.5, 0.0, np.nan, 1e300])
    b = np.array([np.nextafter(1.0, 2.0), 1.193502122111056, np.nextafter(np.nextafter(-3.5, 0), 0), -0.0, np.nan, np.inf])
    d = ulp_diff(a, b)
    assert list(d[:5]) == [1.0, 1.0, 2.0, 0.0, 0.0] and d[5] == np.inf


def _kkt():
    z = np.load(__import__("os").path.join(__import__("casadi_b200.tapeio", fromlist=["GOLDEN_DIR"]).GOLDEN_DIR, "kkt.sym.npz"))
    return {k: np.array(z[k]) for k in z.files}


def dense_sp(nrow, ncol):
    return np.array([nrow, ncol] + [nrow * c for c in range(ncol + 1)] + [r for _ in range(ncol) for r in range(nrow)],
                    np.int64)


@pytest.mark.parametrize("solver", ["ldl", "qr"])
def test_oracle_linsol_matches_reference_mx_solve(solver):
    """oracle_ldl/_solve, oracle_qr/_solve and oracle_mtimes against the reference's own
    MX function [x = solve(K,b,solver); r = mtimes(K,x)-b] mapped serially over 150 KKT systems."""
    sym, case = _kkt(), load_case("kkt_" + solver)
    N, nnz, n = case["N"], int(sym["sp_a"][2 + 60]), 60
    K, B = case["in"][0].reshape(N, nnz), case["in"][1].reshape(N, n)
    X, R = case["out"][0].reshape(N, n), case["out"][1].reshape(N, n)
    for i in range(N):
        if solver == "ldl":
            x = oracle.ldl_factor_solve(sym["sp_a"], sym["sp_lt"], sym["p"], K[i], B[i])[0]
        else:
            x = oracle.qr_factor_solve(sym["sp_a"], sym["sp_v"], sym["sp_r"], sym["prinv"], sym["pc"], K[i], B[i])[0]
        assert_bit_equal(x, X[i], "%s x[%d]" % (solver, i))
        r = oracle.mtimes(K[i], sym["sp_a"], x, dense_sp(n, 1), np.zeros(n), dense_sp(n, 1)) - B[i]
        assert_bit_equal(r, R[i], "%s r[%d]" % (solver, i))
