"""BASELINE config 4: mapaccum (T=100) Monte-Carlo rollouts with reduce_out sums across the GPUs of one node.

  python tools/bench_mc.py --samples 16777216                                  (1 GPU)
  python -m torch.distributed.run --nproc-per-node G ... tools/bench_mc.py --samples ...   (G GPUs, NCCL all-reduce of block sums)

Inputs are a deterministic function of the GLOBAL instance index, so every world size evaluates the same batch and
the printed sums must be bit-identical for 1/2/4/8 GPUs (fixed-shape tree, casadi_b200/csrc/reduce.cu)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casadi_b200 import CudaTape, LAYOUT_SOA, load_tape  # noqa: E402
from casadi_b200.dist import ShardedCudaMap  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", dest="n", type=int, default=1 << 24, help="total Monte-Carlo samples over all GPUs")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t = CudaTape(load_tape("mc"), device=local)
    sm = ShardedCudaMap(t, a.n, reduce_out=[1, 1])
    i = torch.arange(sm.i0, sm.i0 + sm.n, device=dev, dtype=torch.float64)
    # counter-based pseudo noise: a fixed function of (global sample index, component)
    k4 = torch.arange(4, device=dev, dtype=torch.float64)[:, None]
    x0 = torch.sin(0.37 * i[None, :] + k4) * 0.9
    W = torch.empty((200, sm.n), device=dev, dtype=torch.float64)
    for k in range(200):
        W[k] = 0.3 * torch.sin(12.9898 * (i * 1e-3 % 7.0) + 78.233 * k) * torch.cos(0.001 * i + k)
    xs = torch.empty(4, device=dev, dtype=torch.float64)
    js = torch.empty(1, device=dev, dtype=torch.float64)
    arg, res = [x0.data_ptr(), W.data_ptr()], [xs.data_ptr(), js.data_ptr()]
    for _ in range(a.warmup):
        sm.eval_device(arg, res)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        sm.eval_device(arg, res)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        info = t.info()
        v = a.n * a.steps / (float(ms.item()) * 1e-3)
        print(json.dumps({"metric": "mc_rollouts_per_sec", "value": v, "unit": "evals/s", "n_gpus": world, "n": a.n,
                          "steps": a.steps, "ms_per_step": float(ms.item()) / a.steps, "mode": info["mode"],
                          "hbm_GBs_algorithmic": v * info["bytes_in"] / 1e9,
                          "sum_xT_hex": [float(x).hex() for x in xs.cpu()], "sum_J_hex": float(js.cpu()[0]).hex()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
