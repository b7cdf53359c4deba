"""Flat against re-rolled plans (csrc/tape_reroll.hpp) on one B200, device-resident SoA data (run under gpurun):
for every tape the automatic flat plan and the re-rolled variants given in VARIANTS; every variant's outputs must carry
the bits of the flat plan's.  One JSON line per (tape, variant).
usage: sweep_roll.py [tape ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, capi, load_case, load_tape

dev = torch.device("cuda:0")
SIZES = {"cartpole": 1 << 23, "quad": 1 << 21, "quad_fwd": 1 << 20, "quad_jac": 1 << 20, "mc": 1 << 21}
VARIANTS = [("flat", {"CCU_JIT_ROLL": "0"}), ("rolled", {"CCU_JIT_ROLL": "1"}),
            ("rolled, state in the loop scratch", {"CCU_JIT_ROLL": "1", "CCU_JIT_ROLL_REGS": "0"}),
            ("rolled, 1 CTA/SM", {"CCU_JIT_ROLL": "1", "CCU_JIT_MINBLOCKS": "1"})]


def main():
    names = sys.argv[1:] or list(SIZES)
    for name in names:
        case, N = load_case(name), SIZES[name]
        P = case["N"]
        tape = load_tape(name)
        d_in = None
        ref = None
        seen = set()
        for label, env in VARIANTS:
            os.environ.update(env)
            t0 = time.time()
            try:
                t = CudaTape(tape, mode="jit")
            except Exception as e:
                print(json.dumps({"tape": name, "variant": label, "error": str(e)[:300]}), flush=True)
                continue
            finally:
                for k in env:
                    os.environ.pop(k, None)
            i = t.info()
            key = (i["jit_loop_iters"], i["jit_loop_slots"], i["jit_segments"], i["jit_max_regs"], label.endswith("1 CTA/SM"))
            if key in seen:  # the variant did not change the plan (e.g. no loop in this tape)
                t.close()
                continue
            seen.add(key)
            if d_in is None:
                d_in = []
                for a, n in zip(case["in"], t.nnz_in):
                    x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, n)).t().contiguous().to(dev)
                    d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
            d_out = [torch.full((n, N), float("nan"), dtype=torch.float64, device=dev) for n in t.nnz_out]
            best = 1e30
            for r in range(4):
                t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                              layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                if r > 0:
                    best = min(best, t.last_kernel_ms())
            bits = [o.view(torch.int64) for o in d_out]
            if ref is None:
                ref = [b.clone() for b in bits]
                same = True
            else:
                same = all(bool((a == b).all()) for a, b in zip(bits, ref))
            print(json.dumps({"tape": name, "variant": label, "N": N, "ms": best, "evals_s": N / best * 1e3,
                              "frac_fp64": N / best * 1e3 * i["flops"] / 18.46e12,
                              "frac_hbm": N / best * 1e3 * (i["bytes_in"] + i["bytes_out"]) / 6.5488e12,
                              "kernels": i["jit_segments"], "regs": i["jit_max_regs"], "threads": i["jit_threads"],
                              "loop_iters": i["jit_loop_iters"], "loop_body": i["jit_loop_body"], "loop_slots": i["jit_loop_slots"],
                              "tile_slots": i["jit_scratch_slots"], "compile_ms": i["jit_compile_ms"], "same_bits": same,
                              "wall_s": round(time.time() - t0, 1)}), flush=True)
            t.close()
            del d_out
        del d_in
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
