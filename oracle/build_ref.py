#!/usr/bin/env python3
"""Build the UNMODIFIED reference (casadi core + the linsol_ldl / linsol_qr / integrator_rk / rootfinder_newton
plugins) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; the
product (casadi_b200/) never imports, links or executes anything built here.

The reference sources are compiled where they lie under /root/reference with
plain `g++` (no cmake, nothing copied into this repository).  The three files
the reference's cmake would have generated are produced here, straight from
reference inputs, into oracle/_ref/gen/ (git-ignored):

  * casadi/config.h                 <- casadi/config.h.cmake   (version macros)
  * casadi/core/casadi_export.h     visibility macros (cmake GenerateExportHeader)
  * casadi_runtime_str.h            <- casadi/core/runtime/*.hpp stringified the
                                       way casadi/generate_runtime.cmake does
                                       (only used by the C code generator)

Source list  = the `.cpp` entries of CASADI_INTERNAL in
/root/reference/casadi/core/CMakeLists.txt:116-258 (plus fmu2/fmu3: WITH_FMI2/3 default ON,
headers vendored in the reference tree), compile definitions = those of the reference's Release build
(casadi/core/CMakeLists.txt:326-334: CASADI_WITH_THREAD,
CASADI_WITH_THREADSAFE_SYMBOLICS, CASADI_SNPRINTF; -fopenmp -DWITH_OPENMP for
OmpMap, map.cpp:340-386).  No -march flag: like the reference's Release build the
host code has no FMA contraction on x86-64.

Outputs (all under oracle/_ref/, git-ignored, shipped to the GPU box by gpurun):
  lib/libcasadi.so  lib/libcasadi_linsol_ldl.so  lib/libcasadi_linsol_qr.so
  lib/libcasadi_integrator_rk.so  lib/libcasadi_rootfinder_newton.so
  obj/*.o           (kept so tests/integration can relink a patched map.o)
"""
import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CASADI_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CXX = os.environ.get("CXX_REF", "/usr/bin/g++")

DEFINES = [
    "-DCASADI_DEFAULT_COMPILER_PLUGIN=shell", "-DCASADI_IS_RELEASE=0",
    "-DCASADI_MAJOR_VERSION=3", "-DCASADI_MINOR_VERSION=7", "-DCASADI_PATCH_VERSION=2",
    "-DCASADI_SNPRINTF=snprintf", "-DCASADI_VERSION=31", "-DCASADI_WITH_THREAD",
    "-DCASADI_WITH_THREADSAFE_SYMBOLICS", "-DHAVE_MKSTEMPS", "-DUSE_CXX11", "-DWITH_DEEPBIND",
    "-DWITH_DEPRECATED_FEATURES", "-DWITH_DL", "-D_USE_MATH_DEFINES", "-DWITH_FMI2", "-DWITH_FMI3",
]
FMI_INC = ["-I" + os.path.join(REF, "external_packages/FMI-Standard-2.0.2/headers"),
           "-I" + os.path.join(REF, "external_packages/FMI-Standard-3.0/headers")]
FLAGS = ["-std=c++17", "-fopenmp", "-DWITH_OPENMP", "-pthread", "-fPIC", "-O3", "-DNDEBUG",
         "-fvisibility=hidden", "-fvisibility-inlines-hidden", "-w"]


# plugin -> its two source files under casadi/solvers (the plugin class and its registration)
PLUGINS = {
    "linsol_ldl": ("linsol_ldl.cpp", "linsol_ldl_meta.cpp"),
    "linsol_qr": ("linsol_qr.cpp", "linsol_qr_meta.cpp"),
    "linsol_tridiag": ("linsol_tridiag.cpp", "linsol_tridiag_meta.cpp"),
    "linsol_lsqr": ("lsqr.cpp", "lsqr_meta.cpp"),   # (BSplineInterpolant::init fits its coefficients with it)
    "interpolant_linear": ("linear_interpolant.cpp", "linear_interpolant_meta.cpp"),
    "interpolant_bspline": ("bspline_interpolant.cpp", "bspline_interpolant_meta.cpp"),
    "integrator_rk": ("runge_kutta.cpp", "runge_kutta_meta.cpp"),
    "rootfinder_newton": ("newton.cpp", "newton_meta.cpp"),
    "rootfinder_fast_newton": ("fast_newton.cpp", "fast_newton_meta.cpp"),
}


def public_flags():
    """Flags any translation unit that includes the reference's internal headers must use
    (class layouts depend on CASADI_WITH_THREAD, function_internal.hpp:269-272)."""
    return ["-std=c++17", "-DCASADI_SNPRINTF=snprintf", "-DCASADI_WITH_THREAD",
            "-DCASADI_WITH_THREADSAFE_SYMBOLICS", "-DCASADI_VERSION=31", "-DWITH_OPENMP", "-fopenmp",
            "-pthread", "-I" + REF, "-I" + os.path.join(OUT, "gen"),
            "-I" + os.path.join(OUT, "gen", "runtime")]


def core_sources():
    txt = open(os.path.join(REF, "casadi/core/CMakeLists.txt")).read()
    blk = txt[txt.index("set(CASADI_INTERNAL"):]
    blk = blk[:blk.index("\n)\n")]
    blk = re.sub(r"#.*", "", blk)
    srcs = [t for t in re.findall(r"[\w/\.\$\{\}]+\.cpp", blk) if "$" not in t]
    srcs += ["fmu2.cpp", "fmu3.cpp"]  # FMU2_SRC / FMU3_SRC (WITH_FMI2/3 are ON by default)
    seen, out = set(), []
    for s in srcs:
        if s in seen:
            continue
        seen.add(s)
        out.append(os.path.join(REF, "casadi/core", s))
    return out


def runtime_sources():
    txt = open(os.path.join(REF, "casadi/core/runtime/CMakeLists.txt")).read()
    blk = txt[txt.index("set(RUNTIME_SRC"):]
    blk = blk[:blk.index(")")]
    names = re.findall(r"[\w]+\.hpp", blk)
    names += ["casadi_to_mex.hpp", "casadi_from_mex.hpp", "casadi_fmu.hpp"]
    return [os.path.join(REF, "casadi/core/runtime", n) for n in names]


def write_if_changed(path, content):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path) and open(path).read() == content:
        return
    with open(path, "w") as f:
        f.write(content)


EXPORT_H = """#ifndef {G}_H
#define {G}_H
#define {G} __attribute__((visibility("default")))
#define {N}_NO_EXPORT __attribute__((visibility("hidden")))
#define {N}_DEPRECATED __attribute__((__deprecated__))
#define {N}_DEPRECATED_EXPORT {G} {N}_DEPRECATED
#define {N}_DEPRECATED_NO_EXPORT {N}_NO_EXPORT {N}_DEPRECATED
#endif
"""


def generate_headers():
    gen = os.path.join(OUT, "gen")
    # config.h from the reference's template
    tpl = open(os.path.join(REF, "casadi/config.h.cmake")).read()
    sub = {
        "CASADI_MAJOR_VERSION": "3", "CASADI_MINOR_VERSION": "7", "CASADI_PATCH_VERSION": "2",
        "CASADI_IS_RELEASE": "0", "CASADI_VERSION": "3.7.2", "git_revision": "reference-tree",
        "git_describe": "3.7.2", "feature_list": "\\n * dynamic-loading\\n * openmp\\n * thread",
        "CMAKE_BUILD_TYPE": "Release", "CMAKE_CXX_COMPILER_ID": "GNU",
        "CASADI_CMAKE_CXX_COMPILER": CXX, "CASADI_MODULES": "casadi;" + ";".join("casadi_" + p for p in PLUGINS),
        "CASADI_PLUGINS": "Linsol::ldl;Linsol::qr;Linsol::tridiag;Linsol::lsqr;Interpolant::linear;Interpolant::bspline;Integrator::rk;Rootfinder::newton;Rootfinder::fast_newton", "CASADI_INSTALL_PREFIX": os.path.join(OUT),
        "CMAKE_SHARED_LIBRARY_PREFIX": "lib", "CMAKE_SHARED_LIBRARY_SUFFIX": ".so",
        "CMAKE_C_OUTPUT_EXTENSION": ".o", "casadi_lapack_libraries": "",
    }
    tpl = tpl.replace("${CMAKE_CXX_FLAGS} ${CMAKE_CXX_FLAGS_${UPPER_CMAKE_BUILD_TYPE}} "
                      "${EXTRA_CXX_FLAGS_FROM_DEFS}", " ".join(FLAGS))
    tpl = re.sub(r"\$\{(\w+)\}", lambda m: sub.get(m.group(1), ""), tpl)
    write_if_changed(os.path.join(gen, "casadi/config.h"), tpl)
    write_if_changed(os.path.join(gen, "casadi/core/casadi_export.h"),
                     EXPORT_H.format(G="CASADI_EXPORT", N="CASADI"))
    for p in PLUGINS:
        write_if_changed(os.path.join(gen, "casadi/solvers/casadi_%s_export.h" % p),
                         EXPORT_H.format(G="CASADI_%s_EXPORT" % p.upper(), N="CASADI_%s" % p.upper()))
    # runtime strings (what casadi/generate_runtime.cmake emits)
    parts = []
    for f in runtime_sources():
        name = os.path.basename(f).split(".")[0]
        lines = open(f).read().replace("\r", "").split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        body = "".join('\n  "%s\\n"' % ln.replace("\\", "\\\\").replace('"', '\\"')
                       for ln in lines if ln != "" or True)
        parts.append("const char* %s_str =%s;\n\n" % (name, body))
    write_if_changed(os.path.join(gen, "runtime/casadi_runtime_str.h"), "".join(parts))


def compile_one(job):
    src, obj, extra = job
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj, 0, ""
    cmd = [CXX] + FLAGS + DEFINES + extra + FMI_INC + [
        "-I" + REF, "-I" + os.path.join(OUT, "gen"), "-I" + os.path.join(OUT, "gen", "runtime"),
        "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stderr[-4000:]


def build(jobs=None, verbose=True):
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present; oracle/_ref must be prebuilt" % REF)
    generate_headers()
    objdir = os.path.join(OUT, "obj")
    libdir = os.path.join(OUT, "lib")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(libdir, exist_ok=True)
    work = []
    core_objs = []
    for s in core_sources():
        o = os.path.join(objdir, os.path.basename(s)[:-4] + ".o")
        core_objs.append(o)
        work.append((s, o, ["-Dcasadi_EXPORTS"]))
    plug_objs = {}
    for p in PLUGINS:
        plug_objs[p] = []
        for s in PLUGINS[p]:
            o = os.path.join(objdir, "plugin_" + s[:-4] + ".o")
            plug_objs[p].append(o)
            work.append((os.path.join(REF, "casadi/solvers", s), o, ["-Dcasadi_%s_EXPORTS" % p]))
    jobs = jobs or os.cpu_count() or 4
    failed = False
    with cf.ThreadPoolExecutor(jobs) as ex:
        for obj, rc, err in ex.map(compile_one, work):
            if rc != 0:
                failed = True
                sys.stderr.write("FAILED %s\n%s\n" % (obj, err))
    if failed:
        raise RuntimeError("reference build failed")
    lib = os.path.join(libdir, "libcasadi.so")
    if not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in core_objs):
        subprocess.check_call([CXX, "-shared", "-fopenmp", "-pthread", "-o", lib] + core_objs + ["-ldl"])
    for p, objs in plug_objs.items():
        plib = os.path.join(libdir, "libcasadi_%s.so" % p)
        if not os.path.exists(plib) or any(os.path.getmtime(o) > os.path.getmtime(plib) for o in objs):
            subprocess.check_call([CXX, "-shared", "-fopenmp", "-pthread", "-o", plib] + objs +
                                  ["-L" + libdir, "-lcasadi", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        print("reference built:", lib)
    return lib


if __name__ == "__main__":
    build()
