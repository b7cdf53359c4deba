"""Sweep of specialisation knobs given as environment settings, one fresh tape per setting (run under gpurun):
device-resident SoA data, best of 3 after a warm-up, every setting must reproduce the bits of the first one.
usage: sweep_env.py <budget_s> <tape>[@N][,<tape>...] "<K=V K=V ...>" ["<K=V ...>" ...]      ("" = the automatic plan)
  tape: a golden tape name, or "kkt" = BASELINE config 5 as CudaMap lowers it (bench.kkt_tape: LDL + solve + residual)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from casadi_b200 import CudaTape, LAYOUT_SOA, capi, load_case, load_tape  # noqa: E402

T0 = time.time()
SIZES = {"cartpole": 1 << 23, "quad": 1 << 21, "quad_fwd": 1 << 20, "quad_adj": 1 << 20, "quad_jac": 1 << 20,
         "rocket_hess": 1 << 19, "mc": 1 << 21, "kkt": 1 << 20}


def make(name):
    if name == "kkt":
        import bench
        return bench.kkt_tape(0, "jit"), load_case("kkt_ldl")
    return CudaTape(load_tape(name), mode="jit"), load_case(name)


def main():
    budget = float(sys.argv[1])
    for spec in sys.argv[2].split(","):  # several tapes share one process (one import of torch)
        sweep(budget, spec, sys.argv[3:] or [""])


def sweep(budget, spec, settings):
    name, _, n = spec.partition("@")
    N = int(n) if n else SIZES[name]
    dev = torch.device("cuda:0")
    d_in = d_out = None
    ref = None
    for s in settings:
        if time.time() - T0 > budget:
            print(json.dumps({"tape": name, "env": s, "skipped": "budget"}), flush=True)
            continue
        env = dict(kv.split("=", 1) for kv in s.split())
        os.environ.update(env)
        t0 = time.time()
        try:
            t, case = make(name)
            if d_in is None:
                P = case["N"]
                d_in = []
                for a, nz in zip(case["in"], t.nnz_in):
                    x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, nz)).t().contiguous().to(dev)
                    d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if nz else x)
                d_out = [torch.empty((nz, N), dtype=torch.float64, device=dev) for nz in t.nnz_out]
            for o in d_out:
                o.fill_(float("nan"))
            best = 1e30
            for r in range(4):
                t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                              layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                if r > 0:
                    best = min(best, t.last_kernel_ms())
            sig = [int(o.view(torch.int64).sum().item()) for o in d_out]
            same = ref is None or sig == ref
            ref = ref or sig
            i = t.info()
            print(json.dumps({"tape": name, "env": s, "N": N, "ms": best, "evals_s": N / best * 1e3,
                              "frac_fp64": N / best * 1e3 * i["flops"] / 18.46e12,
                              "frac_hbm": N / best * 1e3 * (i["bytes_in"] + i["bytes_out"]) / 6.5488e12,
                              "segs": i["jit_segments"], "regs": i["jit_max_regs"], "threads": i["jit_threads"],
                              "xld": i["jit_cross_loads"], "xst": i["jit_cross_stores"], "slots": i["jit_scratch_slots"],
                              "chained": i.get("jit_chained"), "compile_ms": i["jit_compile_ms"], "same_bits": same,
                              "wall_s": round(time.time() - t0, 1)}), flush=True)
            t.close()
        except Exception as e:
            print(json.dumps({"tape": name, "env": s, "error": str(e)[:300]}), flush=True)
        finally:
            for k in env:
                os.environ.pop(k, None)


if __name__ == "__main__":
    main()
