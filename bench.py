#!/usr/bin/env python3
"""bench.py -- measurement of `f.map(N, "cuda")` (BASELINE.json metric: SX Function evals/s, FP64) on the five
BASELINE configs.

Headline (default `--config quad_ms`, BASELINE.json configs[1]): the quadrotor 12-state, 20-step RK4
multiple-shooting integrator F (7 197-instruction SX tape) and its Jacobian F.jacobian() (77 216 instructions, 114
structural nonzeros), both mapped over N = 1e7 instances per GPU.  One "step" evaluates both tapes for all N
instances; one "eval" = one instance through both.  The other configs (`--config cartpole|hess_lag|mc|kkt`) are
measured the same way and, in the default single-GPU run, reported beside the headline in `configs`; under torchrun
the `mc` config (mapaccum rollouts with reduce_out sums merged by NCCL inside libcasadi_cuda.so) rides along so that
the scaling run also covers the one path with a collective in the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]     our arm (libcasadi_cuda.so; device-resident value,
                                                                      end-to-end through the C++ CudaMap plugin)
  python bench.py --impl reference [...]                               reference arm: the UNMODIFIED reference's OpenMP
                                                                      map on this host's cores (oracle/_ref)
For N>1 launch under torchrun (one rank per GPU); shards are independent, "scaling": "weak".  ONE JSON line on rank 0.

Inputs: the reference's golden instances of each config (drawn from the SURVEY 8(d) distributions by
oracle/gen_models.cpp) repeated periodically over the batch, so EVERY instance of EVERY tile of the timed run is
checked: all periods must have the bits of period 0, and period 0 must match the reference golden.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_START = time.time()


def log(msg):
    """progress on stderr (stdout carries the ONE JSON line)"""
    sys.stderr.write("[bench %6.1fs] %s\n" % (time.time() - T_START, msg))
    sys.stderr.flush()


CONFIGS = {
    # name: tapes (golden fixtures), instances per GPU, ref_bench / cuda_bench workload, BASELINE.json config index
    "cartpole": dict(tapes=["cartpole"], N=1_000_000, ref="cartpole", idx=0, exact=False,
                     workload="cart-pole 4-state ODE, 4 RK4 substeps (522-instr SX tape), f.map(1e6)"),
    "quad_ms": dict(tapes=["quad", "quad_jac"], N=10_000_000, ref="quad_ms", idx=1, exact=False,
                    workload="quadrotor12 RK4x20 multiple-shooting map: F (7197-instr tape) + F.jacobian() (77216-instr tape)"),
    "hess_lag": dict(tapes=["rocket_hess"], N=1_000_000, ref="rocket_hess", idx=2, exact=True,
                     workload="hess_lag of the rocket-landing OCP NLP (21175-instr tape, 391 in / 601 out nnz), 1e6 scenarios"),
    "mc": dict(tapes=["mc"], N=12_500_992, ref="mc", idx=3, exact=False, reduce_out=[1, 1],
               workload="mapaccum T=100 Monte-Carlo rollouts (4012-instr tape), reduce_out sums of x_T and cost; "
                        "12 500 992 samples per GPU (1.0e8 on 8 GPUs), block sums merged by NCCL"),
    "kkt": dict(tapes=["kkt_ldl"], N=1_000_000, ref="kkt_ldl", idx=4, exact=True,
                workload="MX function [x = solve(K,b,'ldl'); r = K*x-b], KKT n=60 nnz(K)=368, shared sparsity, 1e6 systems"),
}
EVAL = {"quad_ms": "one instance through F and its Jacobian"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="quad_ms", choices=sorted(CONFIGS))
    ap.add_argument("--instances", dest="n", type=int, default=0, help="instances per GPU per step (0 = the config's size)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-interp", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only: do not measure the other configs")
    ap.add_argument("--cpu-n", type=int, default=100_000, help="instances of the CPU baseline sample (>= 1e5, SURVEY 8d)")
    ap.add_argument("--budget-s", type=float, default=420.0,
                    help="wall-clock budget of the whole run: secondary configs / legs that would start after it are skipped (and say so)")
    ap.add_argument("--e2e-gb", type=float, default=6.0,
                    help="host bytes (in + out, GB) the end-to-end leg of a SECONDARY config may allocate: its batch is capped accordingly")
    return ap.parse_args()


def config_dict(name, n):
    """`config` of the JSON line: identical for both arms (the reference arm's bounded sample is in cpu_baseline.sample)."""
    c = CONFIGS[name]
    return {"workload": c["workload"], "name": name, "baseline_config": c["idx"], "instances_per_gpu": n,
            "eval": EVAL.get(name, "one instance through the mapped function")}


# ------------------------------------------------------------------------------------------ reference arm
def ref_bench_exe():
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_bench")
    return exe if os.path.exists(exe) else None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_ref_bench(workload, n, mode, reps, warm, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    out = subprocess.run([ref_bench_exe(), workload, str(n), mode, str(threads), str(reps), str(warm)], env=env,
                         capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        raise RuntimeError("ref_bench rc=%d: %s" % (out.returncode, (out.stderr or out.stdout)[-300:]))
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_oracle_port(name, n, reps, warm):
    """Fallback when oracle/_ref did not travel: the C restatement (oracle/oracle.c), one core."""
    import oracle
    from casadi_b200.tapeio import load_case, load_tape
    import numpy as np
    secs = []
    jobs = []
    for tn in CONFIGS[name]["tapes"]:
        tape, case = load_tape(tn), load_case(tn)
        reps_in = (n + case["N"] - 1) // case["N"]
        jobs.append((tape, [np.tile(a, reps_in)[:n * int(z)] for a, z in zip(case["in"], tape["nnz_in"])]))
    for r in range(reps + warm):
        t0 = time.perf_counter()
        for tape, ins in jobs:
            oracle.map_eval(tape, n, ins)
        if r >= warm:
            secs.append(time.perf_counter() - t0)
    return {"n": n, "threads": 1, "secs_total": sum(secs), "secs_median": statistics.median(secs), "reps": reps}


def cpu_sample(name, n_target, reps, warm, with_serial=True):
    """The reference's own map of the config on this host: OpenMP on all cores (the value) and serial on one core."""
    cores = host_cores()
    wl = CONFIGS[name]["ref"]
    if ref_bench_exe() and not (wl.startswith("kkt") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lib", "libcasadi_linsol_ldl.so"))):
        n = max(cores, (n_target + cores - 1) // cores * cores)
        r = run_ref_bench(wl, n, "openmp", reps, warm, cores)
        out = {"value": r["n"] * r["reps"] / r["secs_total"], "unit": "evals/s", "cores": r["threads"], "kind": "reference",
               "sample": "%d instances x %d reps, reference f.map(n/T,'serial').map(T,'openmp'), T = %d" % (r["n"], r["reps"], r["threads"]),
               "ms_per_step": 1e3 * r["secs_total"] / r["reps"], "n": r["n"]}
        if with_serial:
            s = run_ref_bench(wl, n_target, "serial", 1, 0, 1)
            out["serial"] = {"value": s["n"] * s["reps"] / s["secs_total"], "unit": "evals/s", "cores": 1,
                             "sample": "%d instances x 1 rep, reference f.map(n,'serial')" % s["n"]}
        return out
    n = 2000
    r = run_oracle_port(name, n, reps, warm)
    return {"value": r["n"] * r["reps"] / r["secs_total"], "unit": "evals/s", "cores": 1, "kind": "port",
            "sample": "%d instances x %d reps, oracle/oracle.c serial port (oracle/_ref absent)" % (r["n"], r["reps"]),
            "ms_per_step": 1e3 * r["secs_total"] / r["reps"], "n": r["n"]}


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.config
    n_cfg = args.n or CONFIGS[name]["N"]
    cb = cpu_sample(name, args.cpu_n, args.steps, args.warmup, with_serial=False)
    line = {"impl": "reference", "metric": "sx_function_evals_per_sec", "value": cb["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, n_cfg),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for ts, ln in self.rows:
            if not (t0 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def kkt_tape(device, mode=None):
    """BASELINE config 5 as the C++ CudaMap lowers it (cuda_map.cpp Lowering::call_mx): x = solve(K, b, "ldl") traced over
    the shared pattern (ccu_builder_ldl), r = K*x - b (ccu_builder_mtimes on a zero accumulator, then element-wise SUB)."""
    import numpy as np
    from casadi_b200 import capi
    from casadi_b200.linsol import _BorrowedTape
    from casadi_b200.tapeio import GOLDEN_DIR
    L = capi.lib()
    z = np.load(os.path.join(GOLDEN_DIR, "kkt.sym.npz"))
    ll = lambda a: np.ascontiguousarray(a, np.int64)  # noqa: E731
    sp_a, sp_lt, p = ll(z["sp_a"]), ll(z["sp_lt"]), ll(z["p"])
    n, nnz = int(sp_a[1]), int(sp_a[2 + int(sp_a[1])])
    b = L.ccu_builder_create()
    pa = lambda a: a.ctypes.data_as(capi.c_ll_p)  # noqa: E731
    K = ll([L.ccu_builder_input(b, 0, k) for k in range(nnz)])
    rhs = [L.ccu_builder_input(b, 1, k) for k in range(n)]
    x = ll(rhs)
    capi.check(L.ccu_builder_ldl(b, pa(sp_a), pa(sp_lt), pa(p), pa(K), pa(x), 1, None))
    sp_x = ll([n, 1, 0, n] + list(range(n)))  # dense column
    zero = L.ccu_builder_const(b, 0.0)
    acc = ll([zero] * n)
    capi.check(L.ccu_builder_mtimes(b, pa(K), pa(sp_a), pa(x), pa(sp_x), pa(acc), pa(sp_x)))
    for k in range(n):
        capi.check(L.ccu_builder_output(b, 0, k, int(x[k])))
        capi.check(L.ccu_builder_output(b, 1, k, L.ccu_builder_op(b, 2, int(acc[k]), rhs[k])))  # OP_SUB
    nin, nout = ll([nnz, n]), ll([n, n])
    capi.check(L.ccu_set_default_mode(capi.MODES[mode]))
    try:
        h = L.ccu_builder_finish(b, 2, pa(nin), 2, pa(nout), int(device))
    finally:
        L.ccu_set_default_mode(-1)
        L.ccu_builder_destroy(b)
    if not h:
        raise capi.CcuError(capi.last_error())

    class Owned(_BorrowedTape):
        def close(self):
            if self.handle:
                capi.lib().ccu_tape_destroy(self.handle)
            self.handle = None
    return Owned(h, [nnz, n], [n, n], device)


UNARY_OPS = {0, 5, 6, 7, 10, 11, 12, 13, 14, 15, 16, 17, 18, 23, 26, 27, 29, 30, 33, 36, 37, 38, 39, 40, 41, 42, 86, 93, 94}


def issue_slots(tape):
    """FP64 issue slots the specialised kernels need per evaluation (informational, SURVEY 8d "weighted count"): the tape
    after value numbering, one slot per +,-,*, 9 per IEEE division or sqrt (3 when the divisor is a compile-time constant:
    its reciprocal is hoisted), 23 per sin/cos pair on one operand (14 for a lone sin or cos), 40 per other libm call --
    the instruction counts of the branch-free sequences in csrc/ccu_ops.cuh."""
    op, i0, i1, i2, d = (tape[k] for k in ("op", "i0", "i1", "i2", "d"))
    last, vn, kind = {}, {}, []
    slots = 0
    trig = {}
    for k in range(len(op)):
        o = int(op[k])
        if o == 46:
            continue
        if o == 44:
            key = ("c", float(d[k]).hex())
        elif o == 45:
            key = ("i", int(i1[k]), int(i2[k]))
        else:
            a = last[int(i1[k])]
            key = (o, a, a if o in UNARY_OPS else last[int(i2[k])])
        v = vn.get(key)
        if v is None:
            v = vn[key] = len(kind)
            kind.append(key[0])
            if isinstance(o, int) and o not in (44, 45):
                if o in (4, 36):
                    slots += 3 if (o == 4 and kind[key[2]] == "c") else 9
                elif o == 10:
                    slots += 9
                elif o in (13, 14):
                    trig.setdefault(key[1], set()).add(o)
                elif o in (0, 88):
                    pass
                elif o in (1, 2, 3, 5, 11, 12, 19, 20, 21, 22, 23, 24, 25, 29, 30, 31, 32, 34, 35):
                    slots += 1
                else:
                    slots += 40
        last[int(i0[k])] = v
    for ops in trig.values():
        slots += 23 if len(ops) == 2 else 14
    return slots


def golden_case(name):
    """(inputs, outputs) of the reference golden of one tape of a config; the kkt config adds the residual K*x-b, which
    the golden does not hold (exactly 0 is not expected: it is checked against the oracle's mtimes in tests/)."""
    from casadi_b200 import load_case
    return load_case(name)


def measure_config(name, args, ctx, K, W, headline):
    """Device-resident measurement of one config on this rank; returns the pieces of its JSON entry."""
    import numpy as np
    import torch
    from casadi_b200 import CudaTape, LAYOUT_SOA, capi, load_tape
    from casadi_b200.dist import ShardedCudaMap
    dev, local, world, dist = ctx["dev"], ctx["local"], ctx["world"], ctx["dist"]
    cfg = CONFIGS[name]
    N = args.n if (args.n and headline) else cfg["N"]
    L = capi.lib()
    t_create = time.time()
    tapes = [kkt_tape(local) if tn == "kkt_ldl" else CudaTape(load_tape(tn), device=local) for tn in cfg["tapes"]]
    t_create = time.time() - t_create
    infos = [t.info() for t in tapes]
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    red = cfg.get("reduce_out")
    jobs = []
    for t, tn in zip(tapes, cfg["tapes"]):
        case = golden_case(tn)
        P = case["N"]
        reps = (N + P - 1) // P
        d_in = []
        for a, nz in zip(case["in"], t.nnz_in):
            x = torch.from_numpy(np.ascontiguousarray(a).reshape(P, nz)).t().contiguous().to(dev)
            d_in.append(x.repeat(1, reps)[:, :N].contiguous() if nz else x)
        if red:
            d_out = [torch.empty(nz, dtype=torch.float64, device=dev) for nz in t.nnz_out]
            sm = ShardedCudaMap(t, N * world, reduce_out=red)  # this rank's shard is [rank*N, (rank+1)*N) (N: whole blocks)
            assert sm.n == N, "per-GPU batch must be a multiple of the reduction block for the sharded map"
        else:
            d_out = [torch.empty((nz, N), dtype=torch.float64, device=dev) for nz in t.nnz_out]
            sm = None
        jobs.append(dict(t=t, case=case, P=P, d_in=d_in, d_out=d_out, sm=sm,
                         arg=[x.data_ptr() if x.numel() else None for x in d_in], res=[x.data_ptr() for x in d_out]))

    def run(j):
        if j["sm"] is not None:
            j["sm"].eval_device(j["arg"], j["res"], stream=stream)  # shard kernels + block sums + NCCL all-reduce + tree
        else:
            j["t"].eval_device(N, j["arg"], j["res"], layout=LAYOUT_SOA, stream=sh)

    dom = max(range(len(jobs)), key=lambda k: infos[k]["flops"])

    def step(evs=None):
        for k, j in enumerate(jobs):
            if evs is not None and k == dom:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
            run(j)
            if evs is not None and k == dom:
                b.record(stream)
                evs.append((a, b))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local) if (ctx["rank"] == 0 and headline) else None
    launches0 = L.ccu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = []
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(K):
        step(evs)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = L.ccu_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    dom_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)

    # ---- parity of the timed run's own outputs: every period == period 0 (bits), period 0 == reference golden
    perr = 0.0
    periods = 0
    for j, tn in zip(jobs, cfg["tapes"]):
        P, case = j["P"], j["case"]
        full = N // P
        if red:
            # sums over a periodic batch: compare with the float64 tree-free sum of the golden outputs (tolerance N*eps*sum|x|)
            for o, w in zip(j["d_out"], case["out"]):
                w = w.reshape(P, -1)
                tot = world * (full * w.sum(0) + w[:N - full * P].sum(0))
                bound = world * N * 2.3e-16 * np.abs(w).sum(0) * N / P + 1e-300
                got = o.cpu().numpy()
                if not np.all(np.abs(got - tot) <= bound):
                    raise SystemExit("bench.py: reduce_out sums of %s outside their bound: %r vs %r" % (tn, got, tot))
                perr = max(perr, float(np.max(np.abs(got - tot) / np.maximum(np.abs(tot), 1.0))))
            continue
        for oi, o in enumerate(j["d_out"]):
            bits = o.view(torch.int64)
            first = bits[:, :P]
            body = bits[:, :full * P].reshape(bits.shape[0], full, P)
            if bool((body != first[:, None, :]).any()) or (N > full * P and bool((bits[:, full * P:] != first[:, :N - full * P]).any())):
                raise SystemExit("bench.py: parity check of the timed run failed: %s out%d differs between periods" % (tn, oi))
            periods = full
            if oi < len(case["out"]):
                got = o[:, :P].t().contiguous().cpu().numpy().ravel()
                want = case["out"][oi]
                if cfg["exact"]:
                    if not np.array_equal(got.view(np.uint64), want.view(np.uint64)):
                        raise SystemExit("bench.py: %s out%d is not bit-identical to the reference golden" % (tn, oi))
                else:
                    e = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))
                    perr = max(perr, e)
                    if not e <= 1e-11:
                        raise SystemExit("bench.py: parity check of the timed run failed (%s out%d rel err %g)" % (tn, oi, e))

    # ---- interpreter kernel on the same workload (the path that needs no NVRTC), headline only
    interp = None
    if headline and world == 1 and not args.no_interp and all(i["mode"] == capi.MODE_JIT for i in infos):
        for t in tapes:
            t.set_mode(capi.MODE_INTERP)
        step(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); step(); b.record(stream); torch.cuda.synchronize()
        ims = a.elapsed_time(b)
        ii = tapes[dom].info()
        interp = {"value": N / (ims * 1e-3), "unit": "evals/s", "ms_per_step": ims, "steps": 1,
                  "plan": {k: ii[k] for k in ("threads", "ipt", "slots_shared", "slots_global")}}
        for t in tapes:
            t.set_mode(capi.MODE_JIT)

    # ---- roofline of the dominant tape's kernels
    idom = infos[dom]
    p64, hbm_peak, hbm_src = ctx["p64"], ctx["hbm_peak"], ctx["hbm_src"]
    flops = idom["flops"]
    # algorithmic bytes: mapped inputs + non-reduced outputs, 8 B each
    nbytes = idom["bytes_in"] + (0 if red else idom["bytes_out"])
    t_fp64, t_hbm = flops / (p64 * 1e12), nbytes / (hbm_peak * 1e9)
    bound = "fp64" if t_fp64 >= t_hbm else "hbm"
    ach_f = flops * N / (dom_ms * 1e-3) / 1e12
    ach_b = nbytes * N / (dom_ms * 1e-3) / 1e9
    jit = idom["mode"] == capi.MODE_JIT
    kname = ("ccu_seg x%d per tile (%s tape, specialised)" % (idom["jit_segments"], cfg["tapes"][dom])) if jit \
        else "ccu_interp_kernel (%s tape)" % cfg["tapes"][dom]
    scratch_bytes = 8 * (idom["jit_cross_loads"] + idom["jit_cross_stores"]) if jit else 8 * (idom["spill_loads"] + idom["spill_stores"])
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[cfg["tapes"][dom]]
        if jit and tr["segments"] == idom["jit_segments"] and tr["scratch_slots"] == idom["jit_scratch_slots"]:
            traffic = tr["dram_bytes_per_eval"]
    except Exception:
        pass
    roofline = {"bound": bound, "kernel": kname,
                "achieved": ach_f if bound == "fp64" else ach_b, "peak": p64 if bound == "fp64" else hbm_peak,
                "unit": "TFLOP/s" if bound == "fp64" else "GB/s",
                "frac": (ach_f / p64) if bound == "fp64" else (ach_b / hbm_peak),
                "peak_source": ("FP64 non-FMA issue rate measured live by ccu_fp64_issue_rate (DADD/s, ~3 ms burst at boost clock); "
                                "contraction is off by contract" if bound == "fp64" else hbm_src),
                "traffic": traffic,
                "traffic_note": "DRAM bytes per evaluation (dram__bytes_read.sum + dram__bytes_write.sum over the launches of one tile / "
                                "instances of the tile) from the ncu capture recorded in profiles/r2_traffic.json for this plan; null when "
                                "the plan differs from the profiled one",
                "kernel_ms": dom_ms, "flops_per_eval": flops, "bytes_per_eval": nbytes,
                "fp64": {"achieved": ach_f, "peak": p64, "unit": "TFLOP/s", "frac": ach_f / p64},
                "hbm": {"achieved": ach_b, "peak": hbm_peak, "unit": "GB/s", "frac": ach_b / hbm_peak, "peak_source": hbm_src},
                "roofline_evals_per_s": 1.0 / max(t_fp64, t_hbm),
                "scratch": {"bytes_per_eval": scratch_bytes, "note": "cross-segment work-vector traffic of the plan (loads + stores); not algorithmic bytes"}}
    if cfg["tapes"][dom] != "kkt_ldl":
        try:
            sl = issue_slots(load_tape(cfg["tapes"][dom]))
            roofline["fp64"]["issue_slots_per_eval"] = sl
            roofline["fp64"]["frac_of_issue_rate"] = sl * N / (dom_ms * 1e-3) / 1e12 / p64
            roofline["fp64"]["issue_note"] = ("informational: FP64 issue slots actually needed per evaluation (IEEE division 9, by a constant 3, "
                                               "sin/cos pair 23, after value numbering) x evals/s over the measured issue rate; the headline "
                                               "frac counts ONE flop per tape instruction")
        except Exception:
            pass
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz"):
        # the same fraction against the FP64 rate at the clock the timed region actually ran at
        roofline["fp64"]["frac_at_sustained_clock"] = ach_f / (p64 * clocks["sm_mhz"] / clocks["sm_max_mhz"])
    plan = {}
    for tn, inf in zip(cfg["tapes"], infos):
        plan[tn] = ({k: inf[k] for k in inf if k.startswith("jit_")} if inf["mode"] == capi.MODE_JIT else
                    {k: inf[k] for k in ("threads", "ipt", "slots_shared", "slots_global")})
    io_bytes = sum(i["bytes_in"] + (0 if red else i["bytes_out"]) for i in infos) * N
    res = dict(name=name, N=N, value=world * N * K / (ms * 1e-3), ms_per_step=ms / K, roofline=roofline, launches=int(launches),
               clocks=clocks, interp=interp, perr=perr, periods=periods, plan=plan, tape_create_s=round(t_create, 2),
               mode="jit" if all(i["mode"] == capi.MODE_JIT for i in infos) else "interp", io_gb=io_bytes / 1e9,
               collective=("ccu_comm_allreduce_block_sums (NCCL %d, %d ranks) inside the timed step" % (L.ccu_comm_nccl_version(), world))
               if (red and world > 1) else None)
    if red:
        res["sums_hex"] = [[float(v).hex() for v in o.cpu().numpy()] for o in jobs[0]["d_out"]]
    for j in jobs:
        if j["sm"] is not None:
            j["sm"].close()
    del jobs
    for t in tapes:
        t.close()
    torch.cuda.empty_cache()
    return res


def cuda_bench_exe():
    exe = os.path.join(ROOT, "tests", "integration", "_build", "bin", "cuda_bench")
    return exe if os.path.exists(exe) else None


def e2e_cap(name, gb):
    """largest batch of a secondary config whose caller-side buffers (inputs + outputs) stay within `gb` GB of host memory"""
    from casadi_b200.tapeio import load_tape
    per = 0
    for tn in CONFIGS[name]["tapes"]:
        if tn == "kkt_ldl":
            per += 8 * (368 + 60 + 60 + 60)
            continue
        t = load_tape(tn)
        per += 8 * (int(sum(t["nnz_in"])) + (0 if CONFIGS[name].get("reduce_out") else int(sum(t["nnz_out"]))))
    return max(1024, int(gb * 1e9 / max(per, 1)) // 1024 * 1024)


def measure_e2e(name, N, ctx, reps=2, buffers="pageable"):
    """End to end through the plugin: the C++ CudaMap inside the relinked reference library, ordinary pageable buffers
    (tools/cuda_bench.cpp); every rank runs its own process on its own device."""
    import torch
    world, dist, local, dev = ctx["world"], ctx["dist"], ctx["local"], ctx["dev"]
    cfg = CONFIGS[name]
    exe = cuda_bench_exe()
    if not exe:
        return {"value": None, "unit": "evals/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "tests/integration/_build/bin/cuda_bench missing (built where the reference tree exists)"}
    env = dict(os.environ, CASADI_CUDA_LIB=os.path.join(ROOT, "casadi_b200", "lib", "libcasadi_cuda.so"), CASADI_CUDA_DEVICE=str(local))
    env.pop("CASADI_CUDA_DEVICES", None)
    cmd = [exe, cfg["ref"], str(N), str(reps), "1", buffers] + (["reduce"] if cfg.get("reduce_out") else [])
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    env.pop("CCU_HOST_REGISTER", None)  # (cuda_bench sets it itself for buffers="registered")
    log("e2e %s: %s" % (name, " ".join(cmd[1:])))
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired as e:
        out = subprocess.CompletedProcess(cmd, 124, "", "cuda_bench timed out after 600 s: %s" % ((e.stderr or b"")[-200:],))
    ok = out.returncode == 0
    r = json.loads(out.stdout.strip().splitlines()[-1]) if ok else None
    dt = r["secs_total"] if ok else float("inf")
    if world > 1:
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    if not ok or dt == float("inf"):
        return {"value": None, "unit": "evals/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "cuda_bench failed: %s" % ((out.stderr or out.stdout)[-300:] if not ok else "on another rank")}
    if not r["parity_rel_err"] <= 1e-11:
        raise SystemExit("bench.py: e2e parity check failed (rel err %g)" % r["parity_rel_err"])
    return {"value": world * N * reps / dt, "unit": "evals/s", "steps": reps, "instances_per_gpu": N,
            "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"],
            "construct_s": r["construct_s"], "fstats_last_call_s": r["fstats"], "parity_rel_err": r["parity_rel_err"],
            "buffers": buffers, "registered_buffers": r.get("registered_buffers"),
            "note": ("CudaMap::eval, pageable host buffers: f.map(N,'cuda')(arg,res,iw,w,0) through the reference's public C++ API "
                     "in the relinked libcasadi.so (tools/cuda_bench.cpp); H2D + kernels + D2H and the pinned staging of the pageable "
                     "buffers inside the timed region, host-clock timed; Map construction (construct_s) reported separately")
            if buffers == "pageable" else
            ("the same call with the same malloc buffers, page-locked in place by the library during the warm-up call "
             "(opt-in: CCU_HOST_REGISTER=1 / ccu_host_register, for callers whose buffers outlive the map): no staging copy")}


def main_cuda(args):
    import torch
    import torch.distributed as dist
    from casadi_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    L = capi.lib()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    rate = ctypes.c_double()
    capi.check(L.ccu_fp64_issue_rate(local, ctypes.byref(rate)))
    ctx = dict(rank=rank, world=world, local=local, dev=dev, dist=dist, p64=rate.value / 1e12, hbm_peak=hbm_peak, hbm_src=hbm_src)

    name = args.config
    log("headline %s: device-resident legs" % name)
    head = measure_config(name, args, ctx, K, W, headline=True)
    log("headline value %.4g evals/s, roofline.frac %.3f" % (head["value"], head["roofline"]["frac"]))
    # (several ranks share one host: the caller-side buffers and the reference's repmat'ed Map sparsities of all ranks
    # together stay within ~48 GB, so the per-rank e2e batch shrinks with the world size; one rank runs the full batch)
    e2e = None if args.no_e2e else measure_e2e(name, head["N"] if world == 1 else min(head["N"], e2e_cap(name, 48.0 / world)), ctx)
    cpu = None
    if not args.no_cpu and world == 1 and rank == 0:
        log("headline CPU baseline (reference openmp + serial)")
        try:
            cb = cpu_sample(name, args.cpu_n, 2, 1)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "serial") if k in cb}
        except Exception as e:  # the baseline is a reported number; never lose the GPU line over it
            cpu = {"value": None, "unit": "evals/s", "cores": host_cores(), "kind": "reference", "sample": "failed: %s" % e}
    extras = []
    if not args.no_extra:
        others = [c for c in ("cartpole", "hess_lag", "mc", "kkt", "quad_ms") if c != name] if world == 1 else (["mc"] if name != "mc" else [])
        for c in others:
            over = world == 1 and time.time() - T_START > args.budget_s  # (every rank must take the same decision under torchrun)
            if over:
                extras.append({"config": config_dict(c, CONFIGS[c]["N"]), "value": None,
                               "skipped": "time budget of %.0f s (--budget-s) reached" % args.budget_s})
                continue
            try:
                log("config %s: device-resident legs" % c)
                r = measure_config(c, args, ctx, 3, 3, headline=False)
                entry = {"config": config_dict(c, r["N"]), "value": r["value"], "unit": "evals/s", "ms_per_step": r["ms_per_step"],
                         "steps": 3, "roofline": r["roofline"], "gpu_launches": r["launches"], "parity_rel_err": r["perr"],
                         "periods_checked": r["periods"], "mode": r["mode"], "plan": r["plan"]}
                if r.get("collective"):
                    entry["collective"] = r["collective"]
                if r.get("sums_hex"):
                    entry["sums_hex"] = r["sums_hex"]
                if not args.no_e2e and world == 1 and time.time() - T_START <= args.budget_s:
                    entry["e2e"] = measure_e2e(c, min(r["N"], e2e_cap(c, args.e2e_gb)), ctx)
                if not args.no_cpu and world == 1 and rank == 0 and time.time() - T_START <= args.budget_s:
                    log("config %s: CPU baseline" % c)
                    cb = cpu_sample(c, args.cpu_n, 2, 1)
                    entry["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "serial") if k in cb}
                extras.append(entry)
            except SystemExit:
                raise
            except Exception as e:  # a secondary config never takes the headline down
                extras.append({"config": config_dict(c, CONFIGS[c]["N"]), "value": None, "error": str(e)[:300]})
    if e2e and e2e.get("value") and world == 1 and time.time() - T_START <= args.budget_s:
        # informational last leg (the first to go when the time budget is short): the caller opted in to page-locking of its
        # long-lived buffers (CCU_HOST_REGISTER=1); a bounded batch, e2e.value stays the default staged path at the full size
        try:
            reg = measure_e2e(name, min(head["N"], 2_000_000), ctx, buffers="registered")
            e2e["registered"] = {k: reg.get(k) for k in ("value", "unit", "instances_per_gpu", "registered_buffers", "fstats_last_call_s", "note")}
        except SystemExit:
            raise
        except Exception as e:
            e2e["registered"] = {"value": None, "note": "failed: %s" % str(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cfgd = config_dict(name, head["N"])
    line = {"metric": "sx_function_evals_per_sec", "value": head["value"], "unit": "evals/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic: the reference's golden instances of the config (SURVEY 8d distributions) repeated periodically",
            "config": cfgd,
            "run": {"layout": "SoA device-resident", "l2": "inputs+outputs (%.1f GB) larger than L2" % head["io_gb"],
                    "mode": head["mode"], "tape_create_s": head["tape_create_s"], "plan": head["plan"],
                    "parity": "all %d periods of every output bit-equal to period 0; period 0 vs reference golden rel err %.3g" % (
                        head["periods"], head["perr"])},
            "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": e2e, "interpreter": head["interp"],
            "gpu_launches": head["launches"], "clocks": head["clocks"], "parity_rel_err": head["perr"], "configs": extras}
    if head.get("collective"):
        line["collective"] = head["collective"]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_cuda(a)
