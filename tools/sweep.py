"""Plan sweep of the INTERPRETER kernel on device-resident SoA data (run under gpurun): evals/s per (threads, ipt, S)."""
import json
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_case, load_tape


def run(name, N, plans, reps=3):
    tape, case = load_tape(name), load_case(name)
    t = CudaTape(tape, mode="interp")
    dev = torch.device("cuda:0")
    P = case["N"]
    d_in = []
    for a, n in zip(case["in"], t.nnz_in):
        x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)  # (n, P)
        d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
    d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
    for plan in plans:
        try:
            t.set_plan(*plan)
        except Exception as e:
            print(json.dumps({"tape": name, "plan": plan, "error": str(e)})); continue
        info = t.info()
        best = 1e30
        for r in range(reps + 1):
            t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out], layout=LAYOUT_SOA,
                          stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ms = t.last_kernel_ms()
            if r > 0: best = min(best, ms)
        print(json.dumps({"tape": name, "N": N, "plan": plan, "ms": round(best, 3), "evals_s": N / best * 1e3,
                          "ginstr_s": N * info["n_instr"] / best / 1e6, "S": info["slots_shared"], "G": info["slots_global"],
                          "smem": info["smem_bytes"], "ctas_sm": info["ctas_per_sm"],
                          "spill": info["spill_loads"] + info["spill_stores"]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cartpole", "quad", "quad_jac"]
    if "v3" in which:  # with the min-cut bisection order larger shared windows hold (almost) the whole live set
        run("quad", 1 << 20, [(128, 2, 16), (128, 2, 32), (128, 2, 48), (128, 1, 64), (128, 2, 64), (256, 1, 64), (64, 2, 64), (64, 4, 64)])
        run("quad_jac", 1 << 18, [(128, 2, 16), (128, 2, 24), (128, 2, 32), (128, 2, 48), (128, 1, 64), (128, 2, 64), (128, 1, 96), (128, 2, 96), (64, 2, 96)], reps=2)
        run("rocket_hess", 1 << 18, [(128, 2, 16), (128, 2, 32), (128, 2, 64), (128, 1, 96)], reps=2)
        run("mc", 1 << 20, [(0, 0, 0), (128, 2, 16), (256, 2, 20), (256, 1, 20), (128, 4, 20)], reps=2)
        sys.exit(0)
    if "cartpole" in which:
        run("cartpole", 1 << 22, [(128, 1, 0), (128, 2, 0), (128, 4, 0), (256, 1, 0), (256, 2, 0), (64, 2, 0), (64, 4, 0),
                                  (512, 1, 0), (32, 4, 0), (256, 4, 0), (512, 2, 0), (1024, 1, 0)])
    if "quad" in which:
        run("quad", 1 << 20, [(128, 1, 32), (128, 1, 48), (128, 1, 64), (128, 2, 32), (128, 2, 48), (256, 1, 48),
                              (128, 1, 24), (128, 4, 24), (256, 2, 24), (128, 2, 16), (256, 1, 24), (128, 4, 12)])
    if "quad_jac" in which:
        run("quad_jac", 1 << 18, [(128, 1, 32), (128, 1, 48), (128, 2, 32), (128, 1, 24), (128, 2, 24), (256, 1, 24), (128, 4, 16)], reps=2)
    if "rocket_hess" in which:
        run("rocket_hess", 1 << 18, [(128, 1, 32), (128, 1, 48), (128, 2, 32), (128, 1, 24), (128, 2, 24)], reps=2)
