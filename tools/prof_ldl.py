"""Batched LDL (traced casadi_ldl + casadi_ldl_solve, n=60 KKT pattern) on device-resident SoA data: the command ncu wraps."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaLinsol, LAYOUT_SOA, load_case
from casadi_b200.tapeio import GOLDEN_DIR
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
z = np.load(os.path.join(GOLDEN_DIR, "kkt.sym.npz"))
case = load_case("kkt_ldl")
ls = CudaLinsol("ldl", z["sp_a"], (z["sp_lt"], z["p"]))
t = ls.tape
dev = torch.device("cuda:0")
P = case["N"]
d_in = []
for a, n in zip(case["in"], t.nnz_in):
    x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, n)).t().contiguous().to(dev)
    d_in.append(x.repeat(1, (N + P - 1) // P)[:, :N].contiguous())
d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
for r in range(3):
    t.eval_device(N, [x.data_ptr() for x in d_in], [x.data_ptr() for x in d_out], layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
ms = t.last_kernel_ms()
print(json.dumps({"N": N, "ms": ms, "solves_s": N / ms * 1e3, **{k: v for k, v in t.info().items() if k.startswith("jit")}}))
