"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(ccu_map_eval_host / ccu_map_eval_device), against the reference goldens and the pinned oracle.

Bar (BASELINE.json north_star): bit-exact for +,-,*,/,sqrt, comparisons, if_else and the other exact
operations; within 2 ulp for transcendentals -- the tolerance is ULP_TOL below.
"""
import numpy as np
import pytest

import oracle
from casadi_b200 import CudaMap, CudaTape, LAYOUT_AOS, LAYOUT_SOA, capi, load_case, load_tape
from util import EXACT_OPS, ULP_OPS, assert_bit_equal, exactify, tree_sum, ulp_diff

pytestmark = pytest.mark.gpu
MODES = ["interp", "jit"]  # both product paths: the interpreter kernel and the NVRTC-specialised kernels

ULP_TOL = 2          # transcendentals: |device - reference libm| <= 2 ulp
# whole tapes that chain transcendentals into arithmetic (ulp errors propagate).  Observed on a B200 against the reference's
# goldens (glibc): worst |got - want| / max(|want|, 1) = 3.7e-16 over all composite tapes (profiles/r2_composite_ulp.txt)
COMPOSITE_RTOL = 1e-13


def opcover_output_ops():
    """opcode that produced each output nonzero of the opcover function (single-op outputs only)."""
    t = load_tape("opcover")
    n = len(t["op"])
    producer = {}
    out_op = {}
    depth = {}
    for k in range(n):
        op = int(t["op"][k])
        if op in (44, 45):
            producer[int(t["i0"][k])] = (op, 0)
        elif op == 46:
            out_op[int(t["i2"][k])] = producer[int(t["i1"][k])]
        else:
            d1 = producer[int(t["i1"][k])][1]
            d2 = producer[int(t["i2"][k])][1]
            producer[int(t["i0"][k])] = (op, max(d1, d2) + (0 if op in EXACT_OPS else 1))
    return out_op


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case_name", ["opcover", "opcover_special"])
def test_operator_set_per_op(case_name, mode):
    tape = load_tape("opcover")
    case = load_case(case_name)
    N = case["N"]
    m = CudaMap(tape, N, mode=mode)
    assert m.f.info()["mode"] == capi.MODES[mode]
    got = m(case["in"])[0].reshape(N, -1)
    want = case["out"][0].reshape(N, -1)
    out_op = opcover_output_ops()
    worst = {}
    for j in range(want.shape[1]):
        op, ntrans = out_op[j]
        if ntrans == 0:
            assert_bit_equal(got[:, j], want[:, j], "%s output %d (op %d, exact class)" % (case_name, j, op))
        else:
            u = ulp_diff(got[:, j], want[:, j])
            if op == 86:
                # The reference's erfinv (calculus.hpp:300-327) is not one libm call but two Newton steps
                # y -= (erf(y)-x)/(2/sqrt(pi)*exp(-y*y)): the 1-ulp differences between CUDA's and glibc's
                # erf/exp/log are divided by erf'(y), so the admissible distance is 2 ulp of each of those
                # calls propagated through that quotient (plus 2 ulp of the result itself).
                w = want[:, j]
                with np.errstate(all="ignore"):
                    dydx = 1.0 / (2.0 / np.sqrt(np.pi) * np.exp(-w * w))
                    extra = 6 * np.spacing(1.0) * dydx / np.spacing(np.abs(w))
                u = np.where(np.isfinite(extra), np.maximum(u - extra, 0.0), u)
            worst[j] = (op, float(u.max()))
            # chained transcendentals (e.g. log(fabs(a)+0.1)) still only contain ONE transcendental here
            assert u.max() <= ULP_TOL * ntrans, "%s output %d (op %d): %g ulp at a=%r b=%r c=%r (got %r want %r)" % (
                case_name, j, op, u.max(), case["in"][0][int(u.argmax())], case["in"][1][int(u.argmax())],
                case["in"][2][int(u.argmax())], got[int(u.argmax()), j], want[int(u.argmax()), j])
    print("worst ulp per transcendental output:", worst)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["rocket_hess"])
def test_exact_class_tapes_bit_exact_vs_reference(name, mode):
    tape, case = load_tape(name), load_case(name)
    ops = set(int(o) for o in tape["op"]) - {44, 45, 46}
    assert ops <= EXACT_OPS, ops - EXACT_OPS
    outs = CudaMap(tape, case["N"], mode=mode)(case["in"])
    for j, (g, w) in enumerate(zip(outs, case["out"])):
        assert_bit_equal(g, w, "%s out%d" % (name, j))


@pytest.mark.parametrize("name", ["cartpole", "cartpole1", "quad1", "quad", "quad_fwd", "quad_adj", "quad_jac",
                                  "quad1_jac", "mc", "mcstep", "mapnode"])
@pytest.mark.parametrize("mode", MODES)
def test_exactified_tapes_bit_exact_vs_oracle(name, mode):
    tape, case = exactify(load_tape(name)), load_case(name)
    N = min(case["N"], 200)
    ins = [a[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    want = oracle.map_eval(tape, N, ins)
    got = CudaMap(tape, N, mode=mode)(ins)
    for j, (g, w) in enumerate(zip(got, want)):
        assert_bit_equal(g, w, "%s(exactified) out%d" % (name, j))


@pytest.mark.parametrize("name", ["cartpole", "cartpole1", "quad1", "quad", "quad_fwd", "quad_adj", "quad_jac",
                                  "quad1_jac", "mc", "mcstep", "mapnode"])
@pytest.mark.parametrize("mode", MODES)
def test_composite_tapes_vs_reference(name, mode):
    tape, case = load_tape(name), load_case(name)
    outs = CudaMap(tape, case["N"], mode=mode)(case["in"])
    worst_rel, worst_ulp = 0.0, 0.0
    for j, (g, w) in enumerate(zip(outs, case["out"])):
        scale = np.maximum(np.abs(w), 1.0)
        err = np.abs(g - w) / scale
        assert np.all(np.isfinite(g) == np.isfinite(w))
        assert np.nanmax(err, initial=0.0) <= COMPOSITE_RTOL, "%s out%d: rel err %g" % (name, j, np.nanmax(err))
        worst_rel = max(worst_rel, float(np.nanmax(err, initial=0.0)))
        fin = np.isfinite(w) & (w != 0)
        if fin.any():
            worst_ulp = max(worst_ulp, float(np.max(np.abs(g[fin] - w[fin]) / np.spacing(np.abs(w[fin])))))
    # the observed distance to the reference (glibc transcendentals) -- recorded in DESIGN section 5 (run with -s)
    print("COMPOSITE %s %s: worst rel err %.3g, worst distance %.1f ulp of the reference value" % (name, mode, worst_rel, worst_ulp))


@pytest.mark.parametrize("name,plans", [("cartpole", [(128, 1, 0), (64, 2, 0), (256, 4, 0), (128, 1, 8), (32, 2, 5)]),
                                        ("quad", [(128, 1, 48), (128, 2, 16), (64, 1, 128), (256, 1, 24)]),
                                        ("quad1_jac", [(128, 1, 0), (128, 4, 12), (96, 2, 33)])])
def test_every_plan_gives_identical_bits(name, plans):
    """threads / instances-per-thread / shared-slot budget change where values live, never what is computed."""
    tape, case = load_tape(name), load_case(name)
    t = CudaTape(tape, mode="interp")
    ref = None
    for (threads, ipt, S) in plans:
        t.set_plan(threads, ipt, S)
        outs = CudaMap(t, case["N"])(case["in"])
        if ref is None:
            ref = outs
        else:
            for j, (g, w) in enumerate(zip(outs, ref)):
                assert_bit_equal(g, w, "%s plan %r out%d" % (name, (threads, ipt, S), j))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 127, 128, 129, 255, 257, 1000])
def test_ragged_batch_sizes(N, mode):
    tape, case = load_tape("cartpole1"), load_case("cartpole1")
    N = min(N, case["N"])
    ins = [a[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    t = CudaTape(tape, mode=mode)
    full = CudaMap(t, case["N"])(case["in"])
    got = CudaMap(t, N)(ins)
    for j in range(len(got)):
        assert_bit_equal(got[j], full[j][:N * int(tape["nnz_out"][j])], "N=%d out%d" % (N, j))


@pytest.mark.parametrize("mode", MODES)
def test_null_argument_reads_zero_null_result_skipped(mode):
    # reference semantics: sx_function.cpp:116-117
    tape, case = load_tape("mapnode"), load_case("mapnode")
    N = case["N"]
    ins = list(case["in"])
    m = CudaMap(tape, N, mode=mode)
    zero = m([ins[0], np.zeros_like(ins[1]), ins[2], ins[3]])
    part = m([ins[0], None, ins[2], ins[3]], want=[True, False, True])
    assert part[1] is None
    assert_bit_equal(part[0], zero[0])
    assert_bit_equal(part[2], zero[2])
    want = oracle.map_eval(tape, N, [ins[0], None, ins[2], ins[3]], want=[True, False, True])
    assert np.allclose(part[0], want[0], rtol=1e-13, atol=0)


@pytest.mark.parametrize("mode", MODES)
def test_device_pointer_api_soa_and_aos_layouts(mode):
    import torch
    tape, case = load_tape("quad1"), load_case("quad1")
    N = case["N"]
    t = CudaTape(tape, mode=mode)
    host = CudaMap(t, N)(case["in"])
    dev = torch.device("cuda:0")
    for layout in (LAYOUT_AOS, LAYOUT_SOA):
        d_in, d_out = [], []
        for a, n in zip(case["in"], t.nnz_in):
            x = torch.from_numpy(a.reshape(N, n).copy())
            x = x if layout == LAYOUT_AOS else x.t().contiguous()
            d_in.append(x.to(dev))
        for n in t.nnz_out:
            d_out.append(torch.full((N, n) if layout == LAYOUT_AOS else (n, N), float("nan"), dtype=torch.float64, device=dev))
        torch.cuda.synchronize()
        t.eval_device(N, [x.data_ptr() for x in d_in], [x.data_ptr() for x in d_out], layout=layout,
                      stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for j, y in enumerate(d_out):
            y = y if layout == LAYOUT_AOS else y.t().contiguous()
            assert_bit_equal(y.cpu().numpy().ravel(), host[j], "layout %d out%d" % (layout, j))
        assert t.last_kernel_ms() > 0


@pytest.mark.parametrize("mode", MODES)
def test_reduce_out_fixed_tree_and_reference_sum(mode):
    tape, case, ref = load_tape("mc"), load_case("mc"), load_case("mc_sum")
    N = case["N"]
    t = CudaTape(tape, mode=mode)
    per = CudaMap(t, N)(case["in"])
    red = CudaMap(t, N, reduce_in=[0, 0], reduce_out=[1, 1])(case["in"])
    for j, nnz in enumerate((4, 1)):
        # (1) bit-exact against the documented summation tree applied to the per-instance device results
        assert_bit_equal(red[j], tree_sum(per[j].reshape(N, nnz)), "tree out%d" % j)
        # (2) against the reference's sequential HorzRepsum within the summation-order bound N*eps*sum|x|
        bound = N * np.finfo(float).eps * np.abs(case["out"][j].reshape(N, nnz)).sum(axis=0) + 1e-11 * np.abs(ref["out"][j])
        assert np.all(np.abs(red[j] - ref["out"][j]) <= bound), (red[j], ref["out"][j])


@pytest.mark.parametrize("mode", MODES)
def test_reduce_in_broadcast(mode):
    tape, case = load_tape("quad1"), load_case("quad1")
    N = 300
    u0 = case["in"][1][:4].copy()
    x = np.tile(case["in"][0][:12 * 100], 3)
    t = CudaTape(tape, mode=mode)
    full = CudaMap(t, N)([x, np.tile(u0, N)])
    bc = CudaMap(t, N, reduce_in=[0, 1], reduce_out=[0])([x, u0])
    assert_bit_equal(bc[0], full[0])


@pytest.mark.parametrize("mode", MODES)
def test_large_batch_periodic_inputs_full_size(mode):
    """BASELINE size (N=1e6): inputs repeat with period P, so outputs must repeat bit-for-bit and the first
    period must match the reference golden -- a size-independent property."""
    tape, case = load_tape("cartpole"), load_case("cartpole")
    P, N = case["N"], 1_000_000
    reps = N // P
    ins = [np.tile(a, reps) for a in case["in"]]
    out = CudaMap(tape, P * reps, mode=mode)(ins)[0].reshape(reps, -1)
    assert (out.view(np.uint64) == out[0].view(np.uint64)).all()
    err = np.abs(out[0] - case["out"][0]) / np.maximum(np.abs(case["out"][0]), 1.0)
    assert err.max() <= COMPOSITE_RTOL


def test_zero_batch_is_rejected_like_the_reference():
    from casadi_b200 import CcuError
    with pytest.raises(CcuError):
        CudaMap(load_tape("cartpole1"), 0)


def test_interpreter_and_specialised_kernels_agree_bitwise():
    """Same tape, same inputs: the two product paths must produce identical bits for exact-class tapes, and for
    tapes with transcendentals too (both call the same CUDA math functions on the same operand values)."""
    for name in ("quad", "rocket_hess", "mc"):
        tape, case = load_tape(name), load_case(name)
        a = CudaMap(tape, case["N"], mode="interp")(case["in"])
        b = CudaMap(tape, case["N"], mode="jit")(case["in"])
        for j, (x, y) in enumerate(zip(a, b)):
            assert_bit_equal(x, y, "%s out%d interp vs jit" % (name, j))


@pytest.mark.parametrize("seg,tile", [(200, 0), (1200, 256), (5000, 128), (64, 384)])
def test_jit_segmentation_and_tiling_do_not_change_bits(seg, tile):
    tape, case = load_tape("quad"), load_case("quad")
    t = CudaTape(tape, mode="jit")
    ref = CudaMap(t, case["N"])(case["in"])
    t.set_jit_plan(seg_instr=seg, tile=tile)
    info = t.info()
    assert info["mode"] == capi.MODE_JIT and info["jit_segments"] >= 1
    got = CudaMap(t, case["N"])(case["in"])
    for j, (x, y) in enumerate(zip(got, ref)):
        assert_bit_equal(x, y, "seg=%d tile=%d out%d" % (seg, tile, j))


@pytest.mark.parametrize("env", [
    {"CCU_JIT_SCHED": "0"},                                                  # the reference's instruction order
    {"CCU_JIT_FASTOPS": "0"},                                                # plain div.rn.f64 / sincos() (no speculation)
    {"CCU_JIT_RING": "0", "CCU_JIT_SPILL": "0"},                             # plain batched ld.global live-ins
    {"CCU_JIT_RING": "8", "CCU_JIT_SPILL": "-1", "CCU_JIT_REGVALS": "12"},   # cp.async ring + shared-memory spill rows
    {"CCU_JIT_STAGE": "-1"},                                                 # TMA bulk staging of the live-ins
    {"CCU_JIT_PREFETCH": "1", "CCU_JIT_SBLOCK": "32"},                       # L2 prefetch, warp-blocked scratch
    {"CCU_JIT_CHAIN": "1"},                                                  # all segments linked into one persistent kernel
    {"CCU_JIT_STREAMS": "3"},                                                # tiles in flight on side streams
])
def test_every_specialisation_variant_reproduces_the_reference_bits(env, monkeypatch):
    """Each code-generation / launch variant of the specialised kernels against the reference goldens: bit-exact on
    the exact-class tape (rocket hess_lag), and bit-identical to the default plan on a tape with sin/cos."""
    from util import assert_bit_equal as same
    base = {}
    for name in ("rocket_hess", "quad_adj"):
        tape, case = load_tape(name), load_case(name)
        base[name] = CudaMap(CudaTape(tape, mode="jit"), case["N"])(case["in"])
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name in ("rocket_hess", "quad_adj"):
        tape, case = load_tape(name), load_case(name)
        t = CudaTape(tape, mode="jit")
        t.set_jit_plan(seg_instr=700, tile=256 if "CCU_JIT_STREAMS" in env else 0)
        info = t.info()
        assert info["jit_segments"] > 3
        if "CCU_JIT_CHAIN" in env:
            assert info["jit_chained"] == 1, t.jit_chain_error()
        got = CudaMap(t, case["N"])(case["in"])
        for j, (x, y) in enumerate(zip(got, base[name])):
            same(x, y, "%s %s out%d vs default plan" % (name, env, j))
        if name == "rocket_hess":
            for j, (x, y) in enumerate(zip(got, case["out"])):
                same(x, y[:x.size], "%s %s out%d vs reference" % (name, env, j))


@pytest.mark.parametrize("mode", MODES)
def test_host_path_chunked_pipeline_matches_single_chunk(mode, monkeypatch):
    """ccu_map_eval_host cuts the batch into chunks (H2D / compute / D2H overlap); chunking must not change a bit,
    nor the reduce_out summation tree (chunks are multiples of the 1024-instance reduction block)."""
    tape, case = load_tape("mc"), load_case("mc")
    P, reps = case["N"], 26
    N = P * reps - 37  # ragged last chunk
    ins = [np.tile(a, reps)[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    t = CudaTape(tape, mode=mode)
    monkeypatch.setenv("CCU_HOST_CHUNK", "1048576")
    one = CudaMap(t, N)(ins)
    red_one = CudaMap(t, N, reduce_in=[0, 0], reduce_out=[1, 1])(ins)
    monkeypatch.setenv("CCU_HOST_CHUNK", "1024")
    many = CudaMap(t, N)(ins)
    red_many = CudaMap(t, N, reduce_in=[0, 0], reduce_out=[1, 1])(ins)
    for j in range(2):
        assert_bit_equal(many[j], one[j], "chunked out%d" % j)
        assert_bit_equal(red_many[j], red_one[j], "chunked reduce out%d" % j)
        assert_bit_equal(red_many[j], tree_sum(one[j].reshape(N, -1)), "tree out%d" % j)
    # first period against the reference golden
    err = np.abs(many[1][:P] - case["out"][1]) / np.maximum(np.abs(case["out"][1]), 1.0)
    assert err.max() <= COMPOSITE_RTOL


def test_fast_path_operators_match_the_plain_ones_bitwise():
    """The branch-free division / sin / cos / sincos sequences of the specialised kernels (csrc/ccu_ops.cuh) against
    div.rn.f64 and the CUDA math library on 2^28 generated operands of four classes (raw bit patterns, moderate
    magnitudes, trig arguments across the 2^31 fast-path limit, special values): wherever the fast path does not
    flag its operands the bits must be identical (flagged operands are re-evaluated by the plain operator)."""
    mism, flagged, checks = capi.selftest_fastops(1 << 28, seed=20261018)
    assert checks == 2 * (1 << 28)
    assert mism == 0, "%d of %d fast-path results differ from the plain operators" % (mism, checks)
    assert 0 < flagged < checks  # the generated operands do exercise the flag (specials, denormals, huge arguments)


def test_flagged_threads_re_evaluate_with_the_plain_operators():
    """Operands outside the fast paths' range (zero and denormal numerators, huge angles, inf, nan) in SOME instances:
    those threads run their segment a second time on the plain operators; every instance must still match the
    interpreter kernel (which only uses the plain operators) bit for bit."""
    tape, case = load_tape("quad"), load_case("quad")
    N = case["N"]
    ins = [a.copy() for a in case["in"]]
    x = ins[0].reshape(N, 12)
    u = ins[1].reshape(N, 4)
    x[1::7, :] = 0.0                 # 0 / const: the quotient test of the fast path fails
    u[2::7, :] = 0.0
    x[3::7, 6] = 3.0e9               # |angle| >= 2^31: Payne-Hanek reduction
    x[4::7, 7] = np.inf
    x[5::7, 9:12] = 1e-310           # denormal rates
    x[6::11, 8] = np.nan
    a = CudaMap(tape, N, mode="interp")(ins)
    b = CudaMap(tape, N, mode="jit")(ins)
    for j, (p, q) in enumerate(zip(a, b)):
        assert_bit_equal(p, q, "quad out%d with out-of-range operands: interp vs jit" % j)
    # untouched instances still reproduce the golden
    keep = np.ones(N, bool)
    for r in (1, 2, 3, 4, 5):
        keep[r::7] = False
    keep[6::11] = False
    got = b[0].reshape(N, -1)[keep]
    want = case["out"][0].reshape(N, -1)[keep]
    assert (np.abs(got - want) / np.maximum(np.abs(want), 1.0)).max() <= COMPOSITE_RTOL


# ---- full-size parity of the multi-segment / multi-tile path (device resident, periodic inputs) ------------------
def _periodic_device_eval(name, N, mode="jit"):
    """Evaluates tape `name` for N instances whose inputs repeat the golden case with period P; returns
    (outputs as torch tensors [nnz, N], P, info)."""
    import torch
    tape, case = load_tape(name), load_case(name)
    t = CudaTape(tape, mode=mode)
    dev = torch.device("cuda:0")
    P = case["N"]
    reps = (N + P - 1) // P
    d_in = []
    for a, n in zip(case["in"], t.nnz_in):
        x = torch.from_numpy(a.reshape(P, n)).t().contiguous().to(dev)
        d_in.append(x.repeat(1, reps)[:, :N].contiguous() if n else x)
    d_out = [torch.empty((n, N), dtype=torch.float64, device=dev) for n in t.nnz_out]
    t.eval_device(N, [x.data_ptr() if x.numel() else None for x in d_in], [x.data_ptr() for x in d_out],
                  layout=LAYOUT_SOA, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out, P, case, t.info()


def _assert_every_period_equals_the_first(d_out, P, N, what):
    import torch
    full = N // P
    for j, o in enumerate(d_out):
        bits = o.view(torch.int64)
        first = bits[:, :P]
        body = bits[:, :full * P].reshape(bits.shape[0], full, P)
        bad = (body != first[:, None, :]).any(dim=2).any(dim=0)  # per period
        assert not bool(bad.any()), "%s out%d: period(s) %s differ from period 0" % (
            what, j, torch.nonzero(bad).flatten()[:8].tolist())
        if N > full * P:
            tail = bits[:, full * P:]
            assert bool((tail == first[:, :tail.shape[1]]).all()), "%s out%d: ragged tail differs" % (what, j)


@pytest.mark.parametrize("name,N,exact", [
    ("quad_jac", 1600000 + 77, False),   # 41 segments, 753 scratch slots, >= 2 automatic tiles, ragged last CTA
    ("rocket_hess", 1000000, True),      # BASELINE config 2 at its stated size
    ("quad_adj", 1000003, False),
    ("mc", 10000000, False),             # BASELINE config 3 chunk size (1e7 samples)
])
def test_full_size_multi_tile_parity(name, N, exact):
    """BASELINE-size batches through the specialised kernels: every tile, every CTA and the ragged tail must produce
    the bits of period 0, and period 0 must be the reference golden (bit-exact for the exact-class tape; within the
    composite tolerance where sin/cos are chained into arithmetic)."""
    d_out, P, case, info = _periodic_device_eval(name, N)
    assert info["mode"] == capi.MODE_JIT
    if name == "quad_jac":
        assert info["jit_segments"] > 10 and info["jit_scratch_slots"] > 100
    _assert_every_period_equals_the_first(d_out, P, N, name)
    for j, o in enumerate(d_out):
        got = o[:, :P].t().contiguous().cpu().numpy().ravel()
        want = case["out"][j]
        if exact:
            assert_bit_equal(got, want, "%s out%d period 0 vs reference" % (name, j))
        else:
            err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
            assert err.max() <= COMPOSITE_RTOL, (name, j, err.max())
            print("%s out%d: worst distance to the reference %g ulp" % (name, j, float(ulp_diff(got, want).max())))


def test_full_size_interpreter_matches_specialised_kernels_on_every_tile():
    """The interpreter kernel at 1e6 instances of a spilling tape (global scratch per CTA) against the specialised
    kernels: identical bits everywhere."""
    import torch
    N = 1000000
    a, P, _, ia = _periodic_device_eval("quad", N, mode="interp")
    b, _, _, ib = _periodic_device_eval("quad", N, mode="jit")
    assert ia["mode"] == capi.MODE_INTERP and ib["mode"] == capi.MODE_JIT
    for j, (x, y) in enumerate(zip(a, b)):
        assert bool((x.view(torch.int64) == y.view(torch.int64)).all()), "quad out%d interp vs jit at N=1e6" % j
    _assert_every_period_equals_the_first(a, P, N, "quad (interp)")


def test_host_path_pageable_staging_and_phase_stats(monkeypatch):
    """ccu_map_eval_host on ordinary (pageable) numpy buffers goes through the library's pinned staging; with
    CCU_HOST_STAGING=0 the driver stages the copies itself.  Same bits either way, ragged chunks included, and the
    phase statistics (FStats split h2d / kernel / d2h) are filled in."""
    tape, case = load_tape("quad"), load_case("quad")
    P, reps = case["N"], 1500
    N = P * reps - 13
    ins = [np.tile(a, reps)[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    t = CudaTape(tape, mode="jit")
    monkeypatch.setenv("CCU_HOST_CHUNK", "16384")
    staged = CudaMap(t, N)(ins)
    st = t.last_eval_stats()
    assert st["staged_bytes"] == 8 * N * (sum(t.nnz_in) + sum(t.nnz_out))
    assert st["h2d_ms"] > 0 and st["kernel_ms"] > 0 and st["d2h_ms"] > 0 and st["wall_ms"] > 0 and st["stage_ms"] > 0
    monkeypatch.setenv("CCU_HOST_STAGING", "0")
    direct = CudaMap(t, N)(ins)
    assert t.last_eval_stats()["staged_bytes"] == 0
    for j, (a, b) in enumerate(zip(staged, direct)):
        assert_bit_equal(a, b, "pinned staging vs driver-staged copies, out%d" % j)
    got = staged[0].reshape(N, -1)
    assert_bit_equal(got[P * (reps - 2):P * (reps - 1)].ravel(), got[:P].ravel(), "a late period vs period 0")


def test_host_path_page_locked_caller_buffers(monkeypatch):
    """ccu_host_register / CCU_HOST_REGISTER=1 page-lock the caller's own buffers in place: the host path then copies
    them by DMA directly (nothing is staged), with the bits of the staged evaluation; the tape un-registers what the
    automatic mode registered when it is destroyed."""
    L = capi.lib()
    tape, case = load_tape("quad"), load_case("quad")
    P, reps = case["N"], 1500
    N = P * reps - 13
    ins = [np.tile(a, reps)[:N * int(n)].copy() for a, n in zip(case["in"], tape["nnz_in"])]
    t = CudaTape(tape, mode="jit")
    monkeypatch.setenv("CCU_HOST_CHUNK", "16384")
    staged = CudaMap(t, N)(ins)
    assert t.last_eval_stats()["staged_bytes"] == 8 * N * (sum(t.nnz_in) + sum(t.nnz_out))
    base = L.ccu_host_registered_count()
    # explicit registration of the inputs: only the results (fresh numpy buffers) are still staged
    for a in ins:
        capi.check(L.ccu_host_register(a.ctypes.data, a.nbytes))
    assert L.ccu_host_registered_count() == base + len(ins)
    mixed = CudaMap(t, N)(ins)
    assert t.last_eval_stats()["staged_bytes"] == 8 * N * sum(t.nnz_out)
    for a in ins:
        capi.check(L.ccu_host_unregister(a.ctypes.data))
    assert L.ccu_host_unregister(ins[0].ctypes.data) != 0  # not registered any more: refused
    assert L.ccu_host_registered_count() == base
    # automatic registration of caller-owned inputs AND results (raw C ABI call: the buffers outlive the tape)
    outs = [np.full(N * int(n), np.nan) for n in tape["nnz_out"]]
    monkeypatch.setenv("CCU_HOST_REGISTER", "1")
    t2 = CudaTape(tape, mode="jit")
    a_ptr, r_ptr = capi.ptr_array([a.ctypes.data for a in ins]), capi.ptr_array([o.ctypes.data for o in outs])
    for _ in range(2):
        capi.check(L.ccu_map_eval_host(t2.handle, N, a_ptr, r_ptr))
        assert t2.last_eval_stats()["staged_bytes"] == 0
    assert L.ccu_host_registered_count() == base + len(ins) + len(outs)
    t2.close()
    assert L.ccu_host_registered_count() == base
    monkeypatch.delenv("CCU_HOST_REGISTER")
    for j, (a, b, c) in enumerate(zip(staged, mixed, outs)):
        assert_bit_equal(a, b, "staged vs registered inputs, out%d" % j)
        assert_bit_equal(a, c, "staged vs automatically registered buffers, out%d" % j)
