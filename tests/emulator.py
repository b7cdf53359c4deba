"""Test-side emulator of the device program ISA (casadi_b200/csrc/ccu_isa.h).

Executes the packed words the tape compiler emits, vectorised over the instances with the oracle's
scalar operations, so that the allocator / SPILL-FILL / register-forwarding logic can be checked
bit-for-bit against the reference goldens on a CPU-only box.  TEST INFRASTRUCTURE.
"""
import numpy as np

import oracle

D_END, D_CONST, D_INPUT, D_OUTPUT, D_FILL, D_SPILL = 0, 1, 2, 3, 4, 5
D_BIN_FIRST, D_UN_FIRST = 16, 64
D_NONE = (1 << 18) - 1
F_ACC = (1 << 19) - 1
# device opcode -> reference opcode (calculus.hpp:60-218)
BIN = dict(zip(range(16, 35), [1, 2, 3, 4, 8, 19, 20, 21, 22, 24, 25, 28, 31, 32, 34, 35, 43, 95, 97]))
UN = dict(zip(range(64, 93), [0, 5, 6, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18, 23, 26, 27, 29, 30, 33, 36, 37, 38, 39,
                              40, 41, 42, 86, 93, 94]))


def run_program(words, N, nnz_in, nnz_out, args, slots_shared, slots_global, want=None):
    """args[j]: AoS float64 array or None; returns list of AoS outputs."""
    w = np.full((max(slots_shared, 1), N), np.nan)
    g = np.full((max(slots_global, 1), N), np.nan)
    acc = np.full(N, np.nan)
    ins = [None if a is None else np.asarray(a, np.float64).reshape(N, nnz_in[j]) for j, a in enumerate(args)]
    outs = [None if (want is not None and not want[j]) else np.full((N, nnz_out[j]), np.nan)
            for j in range(len(nnz_out))]
    pc = 0
    words = [int(x) for x in words]

    def src(f):
        return acc if f == F_ACC else w[f]

    while True:
        word = words[pc]; pc += 1
        op = word & 0xff; fd = (word >> 8) & 0x3ffff; fa = (word >> 26) & 0x7ffff; fb = word >> 45
        if op == D_END:
            break
        if op == D_OUTPUT:
            if outs[fd] is not None:
                outs[fd][:, fb] = src(fa)
            continue
        if op == D_SPILL:
            assert fa < slots_global
            g[fa] = src(fb)
            continue
        if op == D_CONST:
            r = np.full(N, np.array([words[pc]], np.uint64).view(np.float64)[0]); pc += 1
        elif op == D_INPUT:
            r = np.zeros(N) if ins[fa] is None else ins[fa][:, fb].copy()
        elif op == D_FILL:
            r = g[fa].copy()
        elif op >= D_UN_FIRST:
            x = src(fa)
            r = oracle.vec_op(UN[op], x, x)
        elif op >= D_BIN_FIRST:
            r = oracle.vec_op(BIN[op], src(fa), src(fb))
        else:
            raise AssertionError("bad opcode %d" % op)
        acc = r
        if fd != D_NONE:
            assert fd < slots_shared, "shared slot %d out of range %d" % (fd, slots_shared)
            w[fd] = r
    return [None if o is None else o.ravel() for o in outs]
