#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the unmodified reference (TEST INFRASTRUCTURE).

Runs here (the container that has /root/reference): builds oracle/_ref via build_ref.py,
compiles oracle/gen_models.cpp against it, runs it, and packs the raw dumps as compressed
.npz fixtures.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import build_ref  # noqa: E402


def read_tape(path):
    b = open(path, "rb").read()
    assert b[:8] == b"CCUTAPE1"
    hdr = np.frombuffer(b, dtype=np.int64, count=4, offset=8)
    n, sz_w, n_in, n_out = (int(v) for v in hdr)
    off = 8 + 32
    nnz_in = np.frombuffer(b, dtype=np.int64, count=n_in, offset=off); off += 8 * n_in
    nnz_out = np.frombuffer(b, dtype=np.int64, count=n_out, offset=off); off += 8 * n_out
    arrs = {}
    for name in ("op", "i0", "i1", "i2"):
        arrs[name] = np.frombuffer(b, dtype=np.int32, count=n, offset=off); off += 4 * n
    arrs["d"] = np.frombuffer(b, dtype=np.float64, count=n, offset=off); off += 8 * n
    assert off == len(b)
    return dict(sz_w=np.int64(sz_w), nnz_in=nnz_in, nnz_out=nnz_out, **arrs)


def read_case(path):
    b = open(path, "rb").read()
    assert b[:8] == b"CCUCASE1"
    N, n_in, n_out = (int(v) for v in np.frombuffer(b, dtype=np.int64, count=3, offset=8))
    off = 8 + 24
    nnz_in = np.frombuffer(b, dtype=np.int64, count=n_in, offset=off); off += 8 * n_in
    nnz_out = np.frombuffer(b, dtype=np.int64, count=n_out, offset=off); off += 8 * n_out
    out = {"N": np.int64(N)}
    for j in range(n_in):
        c = N * int(nnz_in[j])
        out["in%d" % j] = np.frombuffer(b, dtype=np.float64, count=c, offset=off); off += 8 * c
    for j in range(n_out):
        c = N * int(nnz_out[j])
        out["out%d" % j] = np.frombuffer(b, dtype=np.float64, count=c, offset=off); off += 8 * c
    assert off == len(b)
    return out


def build_tool(name, extra=()):
    build_ref.build(verbose=False)
    bindir = os.path.join(build_ref.OUT, "bin")
    os.makedirs(bindir, exist_ok=True)
    exe = os.path.join(bindir, name)
    src = os.path.join(HERE, name + ".cpp")
    deps = [src, os.path.join(HERE, "models.hpp"), os.path.join(ROOT, "tools", "bench_models.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call([build_ref.CXX, "-O2"] + build_ref.public_flags() + list(extra) + [
            src, "-o", exe, "-L" + os.path.join(build_ref.OUT, "lib"), "-lcasadi",
            "-Wl,-rpath,$ORIGIN/../lib"])
    return exe


def main():
    exe = build_tool("gen_models")
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([exe, tmp])
        for fn in sorted(os.listdir(tmp)):
            p = os.path.join(tmp, fn)
            if fn.endswith(".tape"):
                np.savez_compressed(os.path.join(gold, fn + ".npz"), **read_tape(p))
            elif fn.endswith(".case"):
                np.savez_compressed(os.path.join(gold, fn + ".npz"), **read_case(p))
            elif fn.endswith(".sym"):
                b = open(p, "rb").read()
                assert b[:8] == b"CCUSYM01"
                off, arrs = 8, {}
                for name in ("sp_a", "sp_lt", "p", "sp_v", "sp_r", "prinv", "pc"):
                    n = int(np.frombuffer(b, dtype=np.int64, count=1, offset=off)[0]); off += 8
                    arrs[name] = np.frombuffer(b, dtype=np.int64, count=n, offset=off); off += 8 * n
                assert off == len(b)
                np.savez_compressed(os.path.join(gold, fn + ".npz"), **arrs)
            elif fn.endswith(".npz"):
                os.replace(p, os.path.join(gold, fn))
    total = sum(os.path.getsize(os.path.join(gold, f)) for f in os.listdir(gold))
    print("golden fixtures: %d files, %.2f MB" % (len(os.listdir(gold)), total / 1e6))


if __name__ == "__main__":
    main()
