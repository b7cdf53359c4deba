"""Shared helpers for the parity tests."""
import numpy as np

OP_SIN, OP_COS = 13, 14
# opcodes whose device result must be bit-identical to the reference (north_star: +,-,*,/,sqrt,
# comparisons, if_else; plus the other exactly-rounded / exact operations of calculus.hpp)
EXACT_OPS = {0, 1, 2, 3, 4, 5, 10, 11, 12, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 34, 35, 36, 88, 97}
# transcendental opcodes: within 2 ulp of the reference's libm
ULP_OPS = {6, 7, 8, 9, 13, 14, 15, 16, 17, 18, 33, 37, 38, 39, 40, 41, 42, 43, 86, 93, 94, 95}


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def bit_equal_mask(got, want):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    return (bits(got) == bits(want)) | (np.isnan(got) & np.isnan(want))


def assert_bit_equal(got, want, what=""):
    same = bit_equal_mask(got, want)
    if not same.all():
        i = int(np.argmax(~same))
        raise AssertionError("%s: %d/%d values differ (first at %d: %r vs %r)" % (
            what, int((~same).sum()), same.size, i, np.asarray(got).ravel()[i], np.asarray(want).ravel()[i]))


def ulp_diff(got, want):
    """distance in units in the last place (of `want`); 0 where both NaN or bit-equal; inf where one is
    NaN/inf and the other is not the same."""
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    out = np.zeros(got.shape)
    same = bit_equal_mask(got, want) | ((got == 0) & (want == 0))
    fin = np.isfinite(got) & np.isfinite(want) & ~same
    a = got[fin].view(np.int64).copy(); b = want[fin].view(np.int64).copy()
    # map the sign-magnitude bit patterns to a monotone integer line
    a = np.where(a < 0, np.int64(-2**63) - a, a); b = np.where(b < 0, np.int64(-2**63) - b, b)
    # subtract as integers (float64 cannot resolve neighbouring int64 near 2**62); a huge distance between
    # values of opposite sign may wrap, which is harmless here (it stays huge) after the float conversion below
    d = a - b
    out[fin] = np.where(d == np.int64(-2**63), 2.0**63, np.abs(d).astype(np.float64))
    out[~fin & ~same] = np.inf
    return out


def exactify(tape):
    """Derived tape with the transcendental instructions replaced by exactly-rounded ones of the same
    arity (unary -> SQ or TWICE, binary -> MUL or SUB), keeping the slot traffic identical.  Its reference
    result comes from the pinned oracle, which lets the big AD tapes be compared BIT-exactly."""
    t = dict(tape)
    op = np.array(tape["op"], np.int32, copy=True)
    unary = {6: 11, 7: 12, 13: 11, 14: 12, 15: 11, 16: 12, 17: 11, 18: 12, 33: 11, 37: 12, 38: 11, 39: 12, 40: 11,
             41: 12, 42: 11, 86: 12, 93: 11, 94: 12}
    binary = {8: 3, 9: 3, 43: 2, 95: 3}
    for k, v in unary.items():
        op[op == k] = v
    for k, v in binary.items():
        op[op == k] = v
    t["op"] = op
    return t


def tree_sum(x):
    """Reference restatement (numpy) of the fixed-shape summation tree of casadi_b200/csrc/reduce.cu:
    x is (N, nnz); level 0 = balanced pairwise tree inside aligned blocks of 1024 instances (padded with
    +0.0), level 1 = balanced pairwise tree over the block sums (padded to a power of two)."""
    x = np.asarray(x, np.float64)
    N, nnz = x.shape
    B = 1024
    nb = max((N + B - 1) // B, 1)
    pad = np.zeros((nb * B, nnz)); pad[:N] = x
    lvl = pad.reshape(nb, B, nnz)
    while lvl.shape[1] > 1:
        lvl = lvl[:, 0::2, :] + lvl[:, 1::2, :]
    s = lvl[:, 0, :]
    p2 = 1
    while p2 < nb:
        p2 *= 2
    pad2 = np.zeros((p2, nnz)); pad2[:nb] = s
    while pad2.shape[0] > 1:
        pad2 = pad2[0::2] + pad2[1::2]
    return pad2[0]
