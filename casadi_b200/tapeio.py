"""Loading of exported SX tapes and parity cases (.npz written by oracle/make_golden.py).

A tape is the reference's `std::vector<ScalarAtomic> algorithm_`
(casadi/core/sx_function.hpp:37-44,258) in structure-of-arrays form:
op[k], i0[k], i1[k], i2[k], d[k], plus sz_w and the nnz of every input/output.
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_tape(name, directory=None):
    z = np.load(os.path.join(directory or GOLDEN_DIR, name + ".tape.npz"))
    return {k: (int(z[k]) if k == "sz_w" else np.array(z[k])) for k in z.files}


def load_case(name, directory=None):
    z = np.load(os.path.join(directory or GOLDEN_DIR, name + ".case.npz"))
    N = int(z["N"])
    n_in = len([k for k in z.files if k.startswith("in")])
    n_out = len([k for k in z.files if k.startswith("out")])
    return {"N": N, "in": [np.array(z["in%d" % j]) for j in range(n_in)],
            "out": [np.array(z["out%d" % j]) for j in range(n_out)]}
