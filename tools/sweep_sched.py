"""Sweep of the specialisation plan with the reference order (schedule 0) and the min-cut bisection order
(schedule 1) on device-resident SoA data (run under gpurun).  Every plan's outputs are compared bit-for-bit with
the first plan's."""
import json
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_case, load_tape


def run(name, N, plans, reps=2):
    tape, case = load_tape(name), load_case(name)
    t = CudaTape(tape, mode="interp")
    dev = torch.device("cuda:0")
    P = case["N"]
    d_in = []
    for a, n in zip(case["in"], t.nnz_in):
        x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)
        d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
    d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
    ref = None
    for sched, seg, threads, minb in plans:
        try:
            t0 = time.time()
            t.set_jit_schedule(sched)
            t.set_jit_plan(seg, threads, minb, 0)
            cs = time.time() - t0
        except Exception as e:
            print(json.dumps({"tape": name, "plan": [sched, seg, threads, minb], "error": str(e)[:300]}), flush=True)
            continue
        info = t.info()
        for o in d_out:
            o.fill_(float("nan"))
        best = 1e30
        for r in range(reps + 1):
            t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                          layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ms = t.last_kernel_ms()
            if r > 0:
                best = min(best, ms)
        same = None
        if ref is None:
            ref = [o.clone() for o in d_out]
        else:
            same = all(bool((a.view(torch.int64) == b.view(torch.int64)).all()) for a, b in zip(ref, d_out))
        print(json.dumps({"tape": name, "N": N, "sched": sched, "seg": seg, "threads": threads, "minb": minb,
                          "ms": round(best, 3), "evals_s": N / best * 1e3, "gflop_s": N * info["flops"] / best / 1e6,
                          "segs": info["jit_segments"], "slots": info["jit_scratch_slots"], "regs": info["jit_max_regs"],
                          "xld": info["jit_cross_loads"], "xst": info["jit_cross_stores"], "create_s": round(cs, 1),
                          "sched_ms": info["jit_schedule_ms"], "bits_equal_first": same}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["quad", "quad_jac", "rocket_hess", "mc"]
    if "quad" in which:
        run("quad", 1 << 21, [(0, 800, 256, 2), (1, 800, 256, 2), (1, 2000, 256, 2), (1, 4000, 256, 2), (1, 8000, 256, 2),
                              (1, 8000, 256, 1), (1, 2000, 128, 4), (1, 4000, 128, 3)])
    if "quad_jac" in which:
        run("quad_jac", 1 << 20, [(0, 800, 256, 2), (1, 800, 256, 2), (1, 1200, 256, 2), (1, 2000, 256, 2), (1, 4000, 256, 2),
                                  (1, 2000, 256, 1), (1, 4000, 256, 1), (1, 2000, 128, 3)])
    if "rocket_hess" in which:
        run("rocket_hess", 1 << 19, [(0, 800, 256, 2), (1, 800, 256, 2), (1, 2000, 256, 2), (1, 4000, 256, 2), (1, 4000, 256, 1)])
    if "mc" in which:
        run("mc", 1 << 21, [(0, 800, 256, 2), (1, 800, 256, 2), (1, 2000, 256, 2), (1, 4000, 256, 2)])
