"""Offline SASS accounting of the specialised kernels (no GPU needed): every kernel of a tape's plan is compiled for
sm_100a by NVRTC (ccu_tape_jit_compile_check), disassembled with cuobjdump, and its instructions are counted by class.
usage: sass_stats.py <tape>[,<tape>...] [K=V ...]     tape: a golden tape name or "kkt" (bench.kkt_tape: config 5 as lowered)
       env settings K=V select plan knobs (CCU_JIT_IOBASE=0, CCU_JIT_FASTOPS=0, CCU_JIT_SEG=..., ...)
Prints one JSON line per tape: registers, local-memory stack, SASS instructions, FP64 / integer / memory classes, and the
opcode histogram.  Both bodies of a kernel (the branch-free one and its re-evaluation with the plain operators) are in the
counts; CCU_JIT_FASTOPS=0 gives the plain body alone."""
import collections
import ctypes
import glob
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casadi_b200 import CudaTape, capi, load_tape  # noqa: E402


def kernel_stats(cubin):
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", cubin], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            ops[mm.group(1).split(".")[0]] += 1
    return int(m.group(1)), int(m.group(2)), ops


def main():
    env = dict(a.split("=", 1) for a in sys.argv[2:])
    os.environ.update(env)
    os.environ.setdefault("CCU_JIT_CACHE", "off")
    L = capi.lib()
    for name in sys.argv[1].split(","):
        if name == "kkt":
            import bench
            t = bench.kkt_tape(-1)
        else:
            t = CudaTape(load_tape(name), device=-1)
        with tempfile.TemporaryDirectory() as tmp:
            n = L.ccu_tape_jit_compile_check(t.handle, tmp.encode())
            if n < 0:
                print(json.dumps({"tape": name, "error": capi.last_error()}))
                continue
            regs = stack = 0
            ops = collections.Counter()
            files = glob.glob(os.path.join(tmp, "*.cubin"))
            for f in files:
                r, s, o = kernel_stats(f)
                regs, stack = max(regs, r), max(stack, s)
                ops.update(o)
        cls = lambda *names: sum(ops[k] for k in names)  # noqa: E731
        print(json.dumps({"tape": name, "env": env, "kernels": len(files), "max_regs": regs, "max_stack_bytes": stack,
                          "sass": sum(ops.values()), "fp64": cls("DADD", "DMUL", "DFMA", "DSETP"),
                          "integer": cls("IMAD", "IADD3", "LEA", "SHF", "LOP3", "ISETP", "VIADD", "MOV", "UMOV", "LDC", "LDCU"),
                          "ring": cls("LDGSTS", "LDS", "LDGDEPBAR", "DEPBAR"), "global": cls("LDG", "STG", "LD", "ST"),
                          "local": cls("LDL", "STL"), "opcodes": dict(ops.most_common(24))}))


if __name__ == "__main__":
    main()
