#!/usr/bin/env python3
"""Build the integration artefacts for `f.map(N, "cuda")` inside the REAL reference library (TEST INFRASTRUCTURE).

Runs only where the reference tree and its object files exist (this container: oracle/build_ref.py keeps
oracle/_ref/obj/*.o).  Steps:
  1. apply casadi_b200/host/casadi_map_cuda.patch to scratch copies of casadi/core/map.cpp and mapsum.cpp (under /tmp,
     removed afterwards -- no reference source is ever written into this repository);
  2. compile the patched files and the new casadi_b200/host/cuda_map.cpp, cuda_mapsum.cpp with the reference's own flags;
  3. relink libcasadi.so from the reference objects with map.o / mapsum.o replaced and cuda_map.o / cuda_mapsum.o added
     -> tests/integration/_build/lib/ (git-ignored; travels to the GPU box), plugins copied alongside;
  4. build tests/integration/test_cuda_map.cpp against it -> _build/bin/test_cuda_map.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

OUT = os.path.join(HERE, "_build")
HOST = os.path.join(ROOT, "casadi_b200", "host")


def newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=True):
    build_ref.build(verbose=False)
    ref = build_ref.REF
    objdir, libdir, bindir = (os.path.join(OUT, d) for d in ("obj", "lib", "bin"))
    for d in (objdir, libdir, bindir):
        os.makedirs(d, exist_ok=True)
    patch = os.path.join(HOST, "casadi_map_cuda.patch")
    inc = ["-I" + HOST, "-I" + os.path.join(ref, "casadi", "core"), "-I" + ref, "-I" + os.path.join(build_ref.OUT, "gen"),
           "-I" + os.path.join(build_ref.OUT, "gen", "runtime")]
    base = [build_ref.CXX] + build_ref.FLAGS + build_ref.DEFINES + build_ref.FMI_INC + ["-Dcasadi_EXPORTS"] + inc
    map_o, cm_o = os.path.join(objdir, "map.o"), os.path.join(objdir, "cuda_map.o")
    ms_o, cms_o = os.path.join(objdir, "mapsum.o"), os.path.join(objdir, "cuda_mapsum.o")
    hdrs = [os.path.join(HOST, "cuda_map.hpp"), os.path.join(HOST, "cuda_mapsum.hpp"), os.path.join(ROOT, "include", "casadi_cuda.h")]
    if newer(map_o, [patch, os.path.join(ref, "casadi/core/map.cpp")] + hdrs) or \
            newer(ms_o, [patch, os.path.join(ref, "casadi/core/mapsum.cpp")] + hdrs):
        with tempfile.TemporaryDirectory() as tmp:
            core = os.path.join(tmp, "casadi", "core")
            os.makedirs(core)
            for f in ("map.cpp", "mapsum.cpp", "CMakeLists.txt"):
                shutil.copy(os.path.join(ref, "casadi", "core", f), core)
            subprocess.check_call(["patch", "-p1", "-s", "-d", tmp, "-i", patch])
            subprocess.check_call(base + ["-c", os.path.join(core, "map.cpp"), "-o", map_o])
            subprocess.check_call(base + ["-c", os.path.join(core, "mapsum.cpp"), "-o", ms_o])
    if newer(cm_o, [os.path.join(HOST, "cuda_map.cpp")] + hdrs):
        subprocess.check_call(base + ["-c", os.path.join(HOST, "cuda_map.cpp"), "-o", cm_o])
    if newer(cms_o, [os.path.join(HOST, "cuda_mapsum.cpp")] + hdrs):
        subprocess.check_call(base + ["-c", os.path.join(HOST, "cuda_mapsum.cpp"), "-o", cms_o])
    lib = os.path.join(libdir, "libcasadi.so")
    ref_objs = sorted(os.path.join(build_ref.OUT, "obj", f) for f in os.listdir(os.path.join(build_ref.OUT, "obj"))
                      if f.endswith(".o") and not f.startswith("plugin_") and f not in ("map.o", "mapsum.o"))
    new_objs = [map_o, cm_o, ms_o, cms_o]
    if newer(lib, new_objs + ref_objs):
        subprocess.check_call([build_ref.CXX, "-shared", "-fopenmp", "-pthread", "-o", lib] + ref_objs + new_objs + ["-ldl"])
    for p in ("libcasadi_%s.so" % q for q in build_ref.PLUGINS):
        src = os.path.join(build_ref.OUT, "lib", p)
        if newer(os.path.join(libdir, p), [src]):
            shutil.copy(src, libdir)
    exe = os.path.join(bindir, "test_cuda_map")
    src = os.path.join(HERE, "test_cuda_map.cpp")
    if newer(exe, [src, lib, os.path.join(ROOT, "tools", "bench_models.hpp")] + hdrs):
        # (the checker: oracle/_build/liboracle.so evaluates lowered tapes in the host-side checks)
        sys.path.insert(0, ROOT)
        import oracle
        oracle.build_oracle()
        odir = os.path.join(ROOT, "oracle", "_build")
        subprocess.check_call([build_ref.CXX, "-O1", "-g"] + build_ref.public_flags() + ["-I" + os.path.join(ROOT, "oracle"), "-I" + HOST,
                              "-I" + os.path.join(ref, "casadi", "core"),
                              src, "-o", exe, "-L" + libdir, "-lcasadi", "-L" + odir, "-loracle",
                              "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath,$ORIGIN/../../../../oracle/_build"])
    # the plugin benchmark (bench.py's end-to-end leg): same library, the reference's public API only
    bexe = os.path.join(bindir, "cuda_bench")
    bsrc = os.path.join(ROOT, "tools", "cuda_bench.cpp")
    if newer(bexe, [bsrc, lib, os.path.join(ROOT, "tools", "bench_models.hpp")]):
        subprocess.check_call([build_ref.CXX, "-O2"] + build_ref.public_flags() + ["-I" + os.path.join(ROOT, "tools"),
                              bsrc, "-o", bexe, "-L" + libdir, "-lcasadi", "-ldl", "-Wl,-rpath,$ORIGIN/../lib"])
    if verbose:
        print("integration build:", exe)
    return exe


if __name__ == "__main__":
    build()
