"""One tape, one plan, a few evaluations on device-resident SoA data: the command ncu wraps (run under gpurun).
usage: prof_one.py <tape> <schedule> <seg_instr> <threads> <min_blocks> <N> [reps]"""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_case, load_tape

name, sched, seg, threads, minb, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
tape, case = load_tape(name), load_case(name)
t = CudaTape(tape, mode="interp")
if seg <= 0:
    from casadi_b200 import capi
    t.set_mode(capi.MODE_JIT)  # the automatic plan
else:
    t.set_jit_schedule(sched)
    t.set_jit_plan(seg, threads, minb, int(os.environ.get("TILE", "0")))
dev = torch.device("cuda:0")
P = case["N"]
d_in = []
for a, n in zip(case["in"], t.nnz_in):
    x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)
    d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
for r in range(reps):
    t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                  layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms = t.last_kernel_ms()
info = t.info()
print(json.dumps({"tape": name, "N": N, "ms": ms, "evals_s": N / ms * 1e3, **{k: v for k, v in info.items() if k.startswith("jit") or k in ("flops", "mode")}}))
