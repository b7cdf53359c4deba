"""Multi-GPU sharding and the cross-rank reduce_out combine (casadi_b200/dist.py).

CPU: world_size-2 gloo processes exercise the host logic (shard ranges, zero-padded block sums, exact all-reduce,
level-1 tree) with numpy restatements of the two reduction kernels; GPU: the shard entry point of the C ABI."""
import os
import sys

import numpy as np
import pytest

from casadi_b200.dist import BLOCK, shard_range, tree_level1
from util import assert_bit_equal, tree_sum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def block_sums(x):
    """level 0 of the fixed tree (numpy restatement of ccu_block_sums_kernel): x (n, nnz), n starts at a block boundary"""
    n, nnz = x.shape
    nb = (n + BLOCK - 1) // BLOCK
    pad = np.zeros((nb * BLOCK, nnz))
    pad[:n] = x
    lvl = pad.reshape(nb, BLOCK, nnz)
    while lvl.shape[1] > 1:
        lvl = lvl[:, 0::2, :] + lvl[:, 1::2, :]
    return lvl[:, 0, :]


@pytest.mark.parametrize("N,world", [(1, 1), (1000, 2), (1024, 2), (1025, 2), (5000, 4), (100000, 8), (3, 8), (8 * 1024, 8)])
def test_shard_ranges_partition_the_batch_in_whole_blocks(N, world):
    pos = 0
    for r in range(world):
        i0, n = shard_range(N, r, world)
        assert i0 == pos and n >= 0 and (n == 0 or i0 % BLOCK == 0)
        pos += n
    assert pos == N


def _worker(rank, world, port, N, nnz, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from casadi_b200.dist import combine_block_sums
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    x = np.random.default_rng(5).standard_normal((N, nnz)) * np.logspace(-6, 6, nnz)  # same data on every rank
    i0, n = shard_range(N, rank, world)
    nb = (N + BLOCK - 1) // BLOCK
    part = torch.zeros((nb, nnz), dtype=torch.float64)
    if n > 0:
        part[i0 // BLOCK:i0 // BLOCK + (n + BLOCK - 1) // BLOCK] = torch.from_numpy(block_sums(x[i0:i0 + n]))
    combine_block_sums(part)
    q.put((rank, tree_level1(part.numpy())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [5000, 1024 * 7 + 13])
def test_two_rank_gloo_reduce_is_bit_identical_to_single_process(N):
    import torch.multiprocessing as mp
    nnz, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + N) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, nnz, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(5).standard_normal((N, nnz)) * np.logspace(-6, 6, nnz)
    want = tree_sum(x)
    for r in range(world):
        assert_bit_equal(got[r], want, "rank %d" % r)


@pytest.mark.gpu
def test_shard_entry_point_reproduces_single_gpu_reduction_bits():
    """Two 'ranks' emulated on one GPU: shards evaluated separately, block sums added, tree evaluated once."""
    import torch
    from casadi_b200 import CudaMap, CudaTape, LAYOUT_SOA, capi, load_case, load_tape
    from casadi_b200.dist import ShardedCudaMap
    tape, case = load_tape("mc"), load_case("mc")
    P, reps = case["N"], 17
    N = P * reps
    t = CudaTape(tape)
    x0 = np.tile(case["in"][0].reshape(P, 4), (reps, 1))
    W = np.tile(case["in"][1].reshape(P, 200), (reps, 1))
    ref = CudaMap(t, N, reduce_out=[1, 1])([x0.ravel(), W.ravel()])
    dev = torch.device("cuda:0")
    L = capi.lib()
    nb = (N + BLOCK - 1) // BLOCK
    parts = [torch.zeros((nb, 4), dtype=torch.float64, device=dev), torch.zeros((nb, 1), dtype=torch.float64, device=dev)]
    ro = capi.int_array([1, 1])
    for rank in range(2):
        i0, n = shard_range(N, rank, 2)
        a = [torch.from_numpy(x0[i0:i0 + n].T.copy()).to(dev), torch.from_numpy(W[i0:i0 + n].T.copy()).to(dev)]
        mine = [torch.zeros_like(p) for p in parts]
        capi.check(L.ccu_map_eval_shard_device(t.handle, N, i0, n, capi.ptr_array([v.data_ptr() for v in a]),
                                               capi.ptr_array([None, None]), None, ro.ctypes.data_as(capi.c_i_p),
                                               capi.ptr_array([p.data_ptr() for p in mine]), LAYOUT_SOA, None))
        torch.cuda.synchronize()
        for p, m in zip(parts, mine):
            p += m
    outs = [torch.empty(4, dtype=torch.float64, device=dev), torch.empty(1, dtype=torch.float64, device=dev)]
    for p, o, nnz in zip(parts, outs, (4, 1)):
        capi.check(L.ccu_reduce_tree_device(0, p.data_ptr(), N, nnz, o.data_ptr(), None))
    torch.cuda.synchronize()
    for j in range(2):
        assert_bit_equal(outs[j].cpu().numpy(), ref[j], "two-shard reduce out%d" % j)
    # and the world-size-1 path of ShardedCudaMap
    sm = ShardedCudaMap(t, N, reduce_out=[1, 1])
    a = [torch.from_numpy(x0.T.copy()).to(dev), torch.from_numpy(W.T.copy()).to(dev)]
    sm.eval_device([v.data_ptr() for v in a], [o.data_ptr() for o in outs])
    torch.cuda.synchronize()
    for j in range(2):
        assert_bit_equal(outs[j].cpu().numpy(), ref[j], "ShardedCudaMap out%d" % j)


def _n_gpus():
    try:
        from casadi_b200 import capi
        return capi.lib().ccu_device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_single_process_multi_device_map_has_single_gpu_bits():
    """ccu_multi (what CudaMap builds for CASADI_CUDA_DEVICES): the batch sharded over all visible devices of this
    process, reduce_out block sums merged by the library's NCCL all-reduce -- every output bit-identical to the
    single-device evaluation, for plain maps and for reductions (repmat.cpp:127-135 / mapsum.cpp:170-184 semantics)."""
    G = _n_gpus()
    if G < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    from casadi_b200 import CudaMap, CudaMultiMap, CudaTape, load_case, load_tape
    tape, case = load_tape("mc"), load_case("mc")
    P, reps = case["N"], 41
    N = P * reps - 5
    ins = [np.tile(a, reps)[:N * int(n)] for a, n in zip(case["in"], tape["nnz_in"])]
    ins[0][::3] *= -0.0  # signed zeros survive the merge (the all-reduce runs on the bit patterns)
    one = CudaTape(tape, device=0)
    ref_plain = CudaMap(one, N)(ins)
    ref_red = CudaMap(one, N, reduce_out=[1, 1])(ins)
    ref_rin = CudaMap(one, N, reduce_in=[1, 0], reduce_out=[0, 1])([ins[0][:4], ins[1]])
    for devs in ([0, 1], list(range(G))):
        mm = CudaMultiMap(tape, N, devs)
        for j, (a, b) in enumerate(zip(mm(ins), ref_plain)):
            assert_bit_equal(a, b, "plain map on devices %s out%d" % (devs, j))
        mr = CudaMultiMap(tape, N, devs, reduce_out=[1, 1])
        for j, (a, b) in enumerate(zip(mr(ins), ref_red)):
            assert_bit_equal(a, b, "reduce_out on devices %s out%d" % (devs, j))
        mi = CudaMultiMap(tape, N, devs, reduce_in=[1, 0], reduce_out=[0, 1])
        for j, (a, b) in enumerate(zip(mi([ins[0][:4], ins[1]]), ref_rin)):
            assert_bit_equal(a, b, "reduce_in + reduce_out on devices %s out%d" % (devs, j))


@pytest.mark.gpu
def test_cuda_map_plugin_on_all_devices():
    """The C++ CudaMap inside the relinked reference library with CASADI_CUDA_DEVICES=all: the whole integration
    suite (maps, reductions, MapSum, Linsol lowering, derivatives) against the reference's serial map."""
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    import subprocess
    exe = os.path.join(ROOT, "tests", "integration", "_build", "bin", "test_cuda_map")
    assert os.path.exists(exe)
    env = dict(os.environ, CASADI_CUDA_LIB=os.path.join(ROOT, "casadi_b200", "lib", "libcasadi_cuda.so"), CASADI_CUDA_DEVICES="all")
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "integration ok" in r.stdout
