"""Plan sweep in one process (run under gpurun): every BASELINE tape x a list of specialisation plans
(seg_instr, threads, min_blocks), device-resident SoA data.  One JSON line per (tape, plan).
usage: sweep_plans.py [tape ...]        env SWEEP_PLANS="seg:threads:minb,..." overrides the plan list"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaLinsol, CudaTape, LAYOUT_SOA, capi, load_case, load_tape
from casadi_b200.tapeio import GOLDEN_DIR

dev = torch.device("cuda:0")
PLANS = [(0, 0, -1), (1200, 256, 2), (2500, 128, 2), (4400, 128, 2), (7000, 128, 2)]
if os.environ.get("SWEEP_PLANS"):
    PLANS = [tuple(int(v) for v in p.split(":")) for p in os.environ["SWEEP_PLANS"].split(",")]
SIZES = {"cartpole": 1 << 23, "quad": 1 << 21, "quad_fwd": 1 << 20, "quad_adj": 1 << 20, "quad_jac": 1 << 20,
         "rocket_hess": 1 << 19, "mc": 1 << 21, "kkt_ldl": 1 << 20, "kkt_qr": 1 << 19}


def make(name):
    if name.startswith("kkt"):
        z = np.load(os.path.join(GOLDEN_DIR, "kkt.sym.npz"))
        ls = CudaLinsol("ldl", z["sp_a"], (z["sp_lt"], z["p"])) if name == "kkt_ldl" else \
            CudaLinsol("qr", z["sp_a"], (z["sp_v"], z["sp_r"], z["prinv"], z["pc"]))
        return ls, ls.tape
    t = CudaTape(load_tape(name), mode="interp")
    return t, t


def main():
    names = sys.argv[1:] or list(SIZES)
    for name in names:
        keep, t = make(name)
        case, N = load_case(name), SIZES[name]
        P = case["N"]
        d_in = []
        for a, n in zip(case["in"], t.nnz_in):
            x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, n)).t().contiguous().to(dev)
            d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
        d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
        ref = None
        for seg, threads, minb in PLANS:
            t0 = time.time()
            try:
                if seg <= 0:
                    os.environ.pop("CCU_JIT_SEG", None)
                    t.set_jit_plan(0, 0, -1, 0)
                else:
                    t.set_jit_plan(seg, threads, minb, 0)
                t.set_mode(capi.MODE_JIT)
                best = 1e30
                for r in range(4):
                    t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                                  layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
                    torch.cuda.synchronize()
                    if r > 0:
                        best = min(best, t.last_kernel_ms())
                # every plan must produce the bits of the first one
                sig = [int(o.view(torch.int64).sum().item()) for o in d_out]
                same = ref is None or sig == ref
                ref = ref or sig
                i = t.info()
                print(json.dumps({"tape": name, "plan": [seg, threads, minb], "N": N, "ms": best, "evals_s": N / best * 1e3,
                                  "frac_fp64": N / best * 1e3 * i["flops"] / 18.46e12,
                                  "frac_hbm": N / best * 1e3 * (i["bytes_in"] + i["bytes_out"]) / 6.5488e12,
                                  "segs": i["jit_segments"], "regs": i["jit_max_regs"], "threads": i["jit_threads"],
                                  "xld": i["jit_cross_loads"], "xst": i["jit_cross_stores"], "slots": i["jit_scratch_slots"],
                                  "compile_ms": i["jit_compile_ms"], "same_bits": same, "wall_s": round(time.time() - t0, 1)}), flush=True)
            except Exception as e:
                print(json.dumps({"tape": name, "plan": [seg, threads, minb], "error": str(e)[:300]}), flush=True)
        del d_in, d_out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
