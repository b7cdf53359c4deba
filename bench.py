#!/usr/bin/env python3
"""bench.py -- headline measurement for `f.map(N, "cuda")` (BASELINE.json metric: SX Function evals/s, FP64).

Workload (BASELINE.json configs[1]): the quadrotor 12-state, 20-step RK4 multiple-shooting integrator F
(7 197-instruction SX tape) and its Jacobian F.jacobian() (77 216 instructions, 114 structural nonzeros),
both mapped over N = 1e7 instances per GPU.  One "step" evaluates both tapes for all N instances; one
"eval" = one instance through both (so the number is NOT inflated by counting the two tapes separately).
Synthetic inputs of SURVEY 8(d).2: x ~ U(-0.3,0.3)^12, u ~ hover*(1+U(-0.1,0.1))^4.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (libcasadi_cuda.so through its C ABI)
  python bench.py --impl reference [...]                         reference arm: the UNMODIFIED reference's
                                                                 OpenMP map on this host's cores (oracle/_ref)
For N>1 launch under torchrun (one rank per GPU); shards are independent (no data-path collective),
"scaling": "weak".  Prints ONE JSON line on rank 0.
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

WORKLOAD = "quadrotor12 RK4x20 multiple-shooting map: F (7197-instr tape) + F.jacobian() (77216-instr tape)"
HOVER = 1.2 * 9.81 / 4
# measured with ncu on the Jacobian tape's automatic plan (41 segments): 40.0 KB read + 31.6 KB written per evaluation
# (profiles/r1_launches_jac_auto_plan.txt)
NCU_DRAM_BYTES_PER_EVAL = 71569


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--instances", dest="n", type=int, default=10_000_000, help="instances per GPU per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-interp", action="store_true")
    ap.add_argument("--cpu-n-per-core", type=int, default=2048)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ reference arm
def ref_bench_exe():
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_bench")
    return exe if os.path.exists(exe) else None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_ref_bench(n, reps, warm, threads):
    exe = ref_bench_exe()
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads)
    env["OMP_PROC_BIND"] = "false"
    out = subprocess.run([exe, "quad_ms", str(n), "openmp", str(threads), str(reps), str(warm)], env=env,
                         capture_output=True, text=True, timeout=1500)
    if out.returncode != 0:
        raise RuntimeError("ref_bench rc=%d: %s" % (out.returncode, (out.stderr or out.stdout)[-300:]))
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_oracle_port(n, reps, warm):
    """Fallback when oracle/_ref did not travel: the C restatement (oracle/oracle.c), one core."""
    import numpy as np
    import oracle
    from casadi_b200.tapeio import load_tape
    tapes = [load_tape("quad"), load_tape("quad_jac")]
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.3, 0.3, (n, 12)).ravel()
    u = (HOVER * (1 + rng.uniform(-0.1, 0.1, (n, 4)))).ravel()
    secs = []
    for r in range(reps + warm):
        t0 = time.perf_counter()
        oracle.map_eval(tapes[0], n, [x, u])
        oracle.map_eval(tapes[1], n, [x, u, None])
        if r >= warm:
            secs.append(time.perf_counter() - t0)
    return {"n": n, "threads": 1, "secs_total": sum(secs), "secs_median": statistics.median(secs), "reps": reps}


def cpu_sample(args, reps, warm):
    cores = host_cores()
    if ref_bench_exe():
        n = cores * args.cpu_n_per_core
        r = run_ref_bench(n, reps, warm, cores)
        kind = "reference"
    else:
        n = 2000
        r = run_oracle_port(n, reps, warm)
        kind = "port"
    val = r["n"] * r["reps"] / r["secs_total"]
    return {"value": val, "unit": "evals/s", "cores": r["threads"], "kind": kind,
            "sample": "%d instances x %d reps of the same F+Jacobian workload, %s" % (
                r["n"], r["reps"], "reference f.map(n/T,'serial').map(T,'openmp')" if kind == "reference"
                else "oracle/oracle.c serial port"),
            "ms_per_step": 1e3 * r["secs_total"] / r["reps"], "n": r["n"]}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_sample(args, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "sx_function_evals_per_sec", "value": cb["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_step": cb["n"], "eval": "one instance through F and its Jacobian"},
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
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def main_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from casadi_b200 import CudaTape, LAYOUT_SOA, capi, load_case, load_tape
    from casadi_b200.cuda_map import CudaMap

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, K, W = args.n, args.steps, max(args.warmup, 3)

    t_create = time.time()
    tF, tJ = CudaTape(load_tape("quad"), device=local), CudaTape(load_tape("quad_jac"), device=local)
    t_create = time.time() - t_create
    iF, iJ = tF.info(), tJ.info()
    L = capi.lib()
    mode_name = {capi.MODE_INTERP: "interp", capi.MODE_JIT: "jit"}

    # synthetic inputs, resident in HBM, SoA [k][instance] (the coalesced device layout of ccu_map_eval_device)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x = (torch.rand((12, N), generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.6
    u = HOVER * (1 + (torch.rand((4, N), generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.2)
    # the first instances are the reference's golden case, so the timed run is also a parity check
    gold = load_case("quad_jac")
    P = gold["N"]
    x[:, :P] = torch.from_numpy(gold["in"][0].reshape(P, 12).T.copy()).to(dev)
    u[:, :P] = torch.from_numpy(gold["in"][1].reshape(P, 4).T.copy()).to(dev)
    xf = torch.empty((12, N), device=dev, dtype=torch.float64)
    j0 = torch.empty((66, N), device=dev, dtype=torch.float64)
    j1 = torch.empty((48, N), device=dev, dtype=torch.float64)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    argF, resF = [x.data_ptr(), u.data_ptr()], [xf.data_ptr()]
    argJ, resJ = [x.data_ptr(), u.data_ptr(), None], [j0.data_ptr(), j1.data_ptr()]

    def step(evs=None):
        tF.eval_device(N, argF, resF, layout=LAYOUT_SOA, stream=sh)
        if evs is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
        tJ.eval_device(N, argJ, resJ, layout=LAYOUT_SOA, stream=sh)
        if evs is not None:
            b.record(stream)
            evs.append((a, b))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
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
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    jac_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)

    # parity of the timed run's own outputs against the reference golden (first P instances)
    def relerr(got, want):
        return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))
    perr = max(relerr(j0[:, :P].T.contiguous().cpu().numpy().ravel(), gold["out"][0]),
               relerr(j1[:, :P].T.contiguous().cpu().numpy().ravel(), gold["out"][1]))
    if not perr <= 1e-11:
        raise SystemExit("bench.py: parity check of the timed run failed (rel err %g)" % perr)

    # ---- end to end through the host-pointer C-ABI call (what CudaMap::eval does): pinned AoS host buffers,
    # H2D + kernels + D2H inside the timed region
    e2e = None
    Ne = N
    if not args.no_e2e:
        # 11.3 GB of pinned host memory per rank at N = 1e7; when the host cannot pin that much for every rank the
        # end-to-end leg runs on a quarter of the batch (stated in the line) instead of taking the whole run down
        try:
            probe = torch.empty((Ne, 126), dtype=torch.float64, pin_memory=True)
            del probe
            ok = 1
        except Exception:
            ok = 0
        if world > 1:
            t_ok = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            ok = int(t_ok.item())
        if not ok:
            Ne = max(N // 4, 1)
    if not args.no_e2e:
        N_dev, N = N, Ne  # the end-to-end leg below works on N = Ne instances
        hx = torch.empty((N, 12), dtype=torch.float64, pin_memory=True)
        hu = torch.empty((N, 4), dtype=torch.float64, pin_memory=True)
        hx.copy_(x[:, :N].t()); hu.copy_(u[:, :N].t())
        hxf = torch.empty((N, 12), dtype=torch.float64, pin_memory=True)
        hj0 = torch.empty((N, 66), dtype=torch.float64, pin_memory=True)
        hj1 = torch.empty((N, 48), dtype=torch.float64, pin_memory=True)
        pa = lambda *ts: capi.ptr_array([None if t is None else t.data_ptr() for t in ts])  # noqa: E731
        aF, rF = pa(hx, hu), pa(hxf)
        aJ, rJ = pa(hx, hu, None), pa(hj0, hj1)

        def step_host():
            capi.check(L.ccu_map_eval_host(tF.handle, N, aF, rF))
            capi.check(L.ccu_map_eval_host(tJ.handle, N, aJ, rJ))
        Ke = min(K, 3)
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            step_host()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        got = hj0[:P].numpy().ravel()
        if not relerr(got, gold["out"][0]) <= 1e-11:
            raise SystemExit("bench.py: e2e parity check failed")
        e2e = {"value": world * N * Ke / dt, "unit": "evals/s", "steps": Ke,
               "h2d_bytes_per_step": 2 * 16 * 8 * N, "d2h_bytes_per_step": (12 + 114) * 8 * N,
               "note": "ccu_map_eval_host on pinned AoS host buffers (the reference's Map layout): chunked H2D | "
                       "AoS->SoA, tape kernels, SoA->AoS | D2H pipeline inside the timed region, host-clock timed"}
        if N != N_dev:
            e2e["instances_per_gpu"] = N
            e2e["note"] += "; reduced batch: the host could not pin the buffers of the full one"
        del hx, hu, hxf, hj0, hj1
        N = N_dev

    # ---- the same workload on the interpreter kernel (the path that needs no NVRTC), reported beside the headline
    interp = None
    if iJ["mode"] == capi.MODE_JIT and world == 1 and not args.no_interp:
        tF.set_mode(capi.MODE_INTERP); tJ.set_mode(capi.MODE_INTERP)
        step(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); step(); b.record(stream); torch.cuda.synchronize()
        ims = a.elapsed_time(b)
        interp = {"value": N / (ims * 1e-3), "unit": "evals/s", "ms_per_step": ims, "steps": 1,
                  "plan_J": {k: iJ[k] for k in ("threads", "ipt", "slots_shared", "slots_global")}}
        tF.set_mode(capi.MODE_JIT); tJ.set_mode(capi.MODE_JIT)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the interpreter on the Jacobian tape)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    rate = ctypes.c_double()
    capi.check(L.ccu_fp64_issue_rate(local, ctypes.byref(rate)))
    p64 = rate.value / 1e12
    flopsJ, bytesJ = iJ["flops"], iJ["bytes_in"] + iJ["bytes_out"]
    ach = flopsJ * N / (jac_ms * 1e-3) / 1e12
    hbm_ach = bytesJ * N / (jac_ms * 1e-3) / 1e9
    t_fp64, t_hbm = flopsJ / (p64 * 1e12), bytesJ / (hbm_peak * 1e9)
    kname = ("ccu_seg x%d per tile (quad_jac tape, specialised)" % iJ["jit_segments"]) if iJ["mode"] == capi.MODE_JIT \
        else "ccu_interp_kernel (quad_jac tape)"
    scratch_bytes = 8 * (iJ["jit_cross_loads"] + iJ["jit_cross_stores"]) if iJ["mode"] == capi.MODE_JIT \
        else 8 * (iJ["spill_loads"] + iJ["spill_stores"])
    roofline = {"bound": "fp64" if t_fp64 >= t_hbm else "hbm", "kernel": kname,
                "achieved": ach, "peak": p64, "unit": "TFLOP/s", "frac": ach / p64,
                "peak_source": "FP64 non-FMA issue rate measured live by ccu_fp64_issue_rate (DADD/s); contraction is off by contract",
                "traffic": NCU_DRAM_BYTES_PER_EVAL if (iJ["mode"] == capi.MODE_JIT and iJ["jit_segments"] == 41) else None,
                "traffic_note": "DRAM bytes per evaluation (dram__bytes_read.sum + dram__bytes_write.sum summed over the 41 "
                                "ccu_seg launches of one tile / instances of the tile) from the ncu capture "
                                "profiles/r1_launches_jac_auto_plan.txt; algorithmic bytes are "
                                "bytes_per_eval, the rest is the cross-segment work vector (roofline.scratch)",
                "kernel_ms": jac_ms, "flops_per_eval": flopsJ, "bytes_per_eval": bytesJ,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "peak_source": hbm_src},
                "roofline_evals_per_s": 1.0 / max(t_fp64, t_hbm),
                "scratch": {"bytes_per_eval": scratch_bytes, "achieved": scratch_bytes * N / (jac_ms * 1e-3) / 1e9,
                            "unit": "GB/s", "frac_of_hbm_peak": scratch_bytes * N / (jac_ms * 1e-3) / 1e9 / hbm_peak,
                            "note": "cross-segment work-vector traffic through HBM (scratch loads + stores of the plan); not algorithmic bytes"}}
    cpu = None
    if not args.no_cpu and world == 1:
        try:
            cb = cpu_sample(args, 3, 1)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the baseline is a reported number; never lose the GPU line over it
            cpu = {"value": None, "unit": "evals/s", "cores": host_cores(), "kind": "reference", "sample": "failed: %s" % e}
    line = {"metric": "sx_function_evals_per_sec", "value": world * N * K / (ms * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": N, "eval": "one instance through F and its Jacobian",
                       "layout": "SoA device-resident", "l2": "inputs+outputs (%.1f GB) larger than L2" % (
                           (iF["bytes_in"] + iF["bytes_out"] + bytesJ) * N / 1e9),
                       "mode": mode_name[iJ["mode"]], "tape_create_s": round(t_create, 2),
                       "plan_F": {k: iF[k] for k in iF if k.startswith("jit_")} if iF["mode"] == capi.MODE_JIT else
                       {k: iF[k] for k in ("threads", "ipt", "slots_shared", "slots_global")},
                       "plan_J": {k: iJ[k] for k in iJ if k.startswith("jit_")} if iJ["mode"] == capi.MODE_JIT else
                       {k: iJ[k] for k in ("threads", "ipt", "slots_shared", "slots_global")}},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "interpreter": interp, "gpu_launches": int(launches), "clocks": clocks,
            "parity_rel_err": perr}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_cuda(a)
