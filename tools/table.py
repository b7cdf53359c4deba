"""All BASELINE tapes on one B200, device-resident SoA data: specialised kernels (automatic plan) and interpreter
(run under gpurun).  One JSON line per tape: evals/s, nominal FP64 and HBM roofline fractions."""
import ctypes
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaLinsol, CudaTape, LAYOUT_SOA, capi, load_case, load_tape
from casadi_b200.tapeio import GOLDEN_DIR

dev = torch.device("cuda:0")


def time_tape(t, ins_np, N, P, reps=3):
    d_in = []
    for a, n in zip(ins_np, t.nnz_in):
        x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, n)).t().contiguous().to(dev)
        d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
    d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
    best = 1e30
    for r in range(reps + 1):
        t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                      layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, t.last_kernel_ms())
    return best


def report(label, t, ins_np, N, P, p64, hbm):
    info = t.info()
    out = {"tape": label, "N": N, "n_instr": info["n_instr"], "flops": info["flops"], "bytes": info["bytes_in"] + info["bytes_out"]}
    for mode in ("jit", "interp"):
        t.set_mode(capi.MODE_JIT if mode == "jit" else capi.MODE_INTERP)
        n = N if mode == "jit" else max(N // 8, 1 << 16)
        ms = time_tape(t, ins_np, n, P)
        ev = n / ms * 1e3
        out[mode + "_evals_s"] = ev
        if mode == "jit":
            i2 = t.info()
            t_fp64, t_hbm = info["flops"] / p64, out["bytes"] / hbm
            out.update({"bound": "fp64" if t_fp64 > t_hbm else "hbm", "roofline_evals_s": 1 / max(t_fp64, t_hbm),
                        "frac": ev * max(t_fp64, t_hbm), "fp64_frac": ev * t_fp64, "hbm_frac": ev * t_hbm,
                        "segs": i2["jit_segments"], "regs": i2["jit_max_regs"], "threads": i2["jit_threads"],
                        "xld": i2["jit_cross_loads"], "xst": i2["jit_cross_stores"], "slots": i2["jit_scratch_slots"]})
        else:
            out["interp_plan"] = [info["threads"], info["ipt"], info["slots_shared"], info["slots_global"]]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    rate = ctypes.c_double()
    capi.check(capi.lib().ccu_fp64_issue_rate(0, ctypes.byref(rate)))
    p64 = rate.value
    peaks = json.load(open(os.path.join(os.path.dirname(GOLDEN_DIR), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(GOLDEN_DIR), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"] * 1e9
    print(json.dumps({"fp64_issue_rate": p64, "hbm_Bps": hbm}), flush=True)
    only = set(sys.argv[1:])
    for name, N in (("cartpole", 1 << 23), ("quad", 1 << 21), ("quad_fwd", 1 << 20), ("quad_adj", 1 << 20),
                    ("quad_jac", 1 << 20), ("rocket_hess", 1 << 19), ("mc", 1 << 21)):
        if only and name not in only:
            continue
        case = load_case(name)
        t = CudaTape(load_tape(name))
        report(name, t, case["in"], N, case["N"], p64, hbm)
    if only and not (only & {"kkt_ldl", "kkt_qr"}):
        sys.exit(0)
    z = np.load(os.path.join(GOLDEN_DIR, "kkt.sym.npz"))
    case = load_case("kkt_ldl")
    ls = CudaLinsol("ldl", z["sp_a"], (z["sp_lt"], z["p"]))
    report("kkt_ldl n=60 (traced casadi_ldl + solve)", ls.tape, case["in"], 1 << 20, case["N"], p64, hbm)
    case = load_case("kkt_qr")
    ls = CudaLinsol("qr", z["sp_a"], (z["sp_v"], z["sp_r"], z["prinv"], z["pc"]))
    report("kkt_qr n=60 (traced casadi_qr + solve)", ls.tape, case["in"], 1 << 19, case["N"], p64, hbm)
