"""Build casadi_b200/lib/libcasadi_cuda.so (C ABI + sm_100a kernels) with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcasadi_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["interp.cu", "reduce.cu", "capi.cu", "diag.cu"]
CPP_SOURCES = ["tape_compile.cpp"]
# -fmad=false: no FMA contraction, the rounding contract of the whole library (see ccu_ops.cuh)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".hpp", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "casadi_cuda.h"))
    objs = []
    procs = []
    for src in CU_SOURCES + CPP_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            bad = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose:
            print(out)
    if bad:
        raise RuntimeError("libcasadi_cuda build failed")
    if force or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
                              ["-Xcompiler", "-fPIC"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
