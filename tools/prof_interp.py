"""Interpreter kernel on one tape (device-resident SoA): the command ncu wraps.  usage: prof_interp.py <tape> <N>"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_case, load_tape
name, N = sys.argv[1], int(sys.argv[2])
tape, case = load_tape(name), load_case(name)
t = CudaTape(tape, mode="interp")
dev = torch.device("cuda:0")
P = case["N"]
d_in = []
for a, n in zip(case["in"], t.nnz_in):
    x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)
    d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous() if n else x)
d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
for r in range(2):
    t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out], layout=LAYOUT_SOA,
                  stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
ms = t.last_kernel_ms()
i = t.info()
print(json.dumps({"tape": name, "N": N, "ms": ms, "evals_s": N / ms * 1e3, **{k: i[k] for k in ("threads", "ipt", "slots_shared", "slots_global", "spill_loads", "spill_stores", "n_words")}}))
