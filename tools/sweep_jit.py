"""Sweep of the specialised-kernel plan on device-resident SoA data (run under gpurun)."""
import json
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_case, load_tape


def run(name, N, plans, reps=3):
    tape, case = load_tape(name), load_case(name)
    t0 = time.time()
    t = CudaTape(tape, mode="jit")
    print(json.dumps({"tape": name, "create_s": round(time.time() - t0, 2), **{k: v for k, v in t.info().items() if k.startswith("jit")}}), flush=True)
    dev = torch.device("cuda:0")
    P = case["N"]
    d_in = []
    for a, n in zip(case["in"], t.nnz_in):
        x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)
        d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
    d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
    for plan in plans:
        try:
            t0 = time.time()
            if plan is not None:
                t.set_jit_plan(*plan)
            cs = time.time() - t0
        except Exception as e:
            print(json.dumps({"tape": name, "plan": plan, "error": str(e)[:300]})); continue
        info = t.info()
        best = 1e30
        for r in range(reps + 1):
            t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out], layout=LAYOUT_SOA,
                          stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ms = t.last_kernel_ms()
            if r > 0: best = min(best, ms)
        print(json.dumps({"tape": name, "N": N, "plan": plan, "ms": round(best, 3), "evals_s": N / best * 1e3,
                          "gflop_s": N * info["flops"] / best / 1e6, "segs": info["jit_segments"], "slots": info["jit_scratch_slots"],
                          "regs": info["jit_max_regs"], "xld": info["jit_cross_loads"], "xst": info["jit_cross_stores"],
                          "compile_s": round(cs, 1)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cartpole", "quad", "quad_jac", "rocket_hess"]
    # plan = (seg_instr, threads, min_blocks, tile)
    if "cartpole" in which:
        run("cartpole", 1 << 23, [None, (1200, 128, 4, 0), (1200, 256, 2, 0), (1200, 64, 0, 0), (1200, 128, 8, 0)])
    if "quad" in which:
        run("quad", 1 << 21, [None, (600, 128, 0, 0), (2400, 128, 0, 0), (1200, 128, 4, 0), (1200, 128, 0, 1 << 20), (1200, 128, 0, 37888), (1200, 256, 0, 0)])
    if "quad_jac" in which:
        run("quad_jac", 1 << 20, [None, (1200, 128, 4, 0), (2400, 128, 0, 0), (600, 128, 0, 0), (1200, 128, 0, 1 << 18), (1200, 128, 0, 37888)], reps=2)
    if "rocket_hess" in which:
        run("rocket_hess", 1 << 19, [None, (1200, 128, 4, 0), (2400, 128, 0, 0), (1200, 128, 0, 1 << 17)], reps=2)
    if "mc" in which:
        run("mc", 1 << 20, [None, (1200, 128, 4, 0)], reps=2)
