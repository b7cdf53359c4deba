"""Multi-GPU evaluation of a mapped tape: one process per GPU (torch.distributed), contiguous shards of whole
1024-instance reduction blocks, tape replicated, NO data-path collective for plain maps.  Only reduce_out
communicates: every rank writes the level-0 sums of its blocks at their global positions into a zero vector, one
all-reduce merges the disjoint supports exactly, and every rank evaluates the same level-1 tree -- the result is
bit-identical for 1/2/4/8 GPUs (HorzRepsum/MapSum semantics up to the documented summation order,
casadi/core/repmat.cpp:127-135, mapsum.cpp:154-186).

On GPUs the all-reduce is issued by libcasadi_cuda.so itself (ccu_comm_allreduce_block_sums: NCCL over NVLink on the
64-bit patterns, csrc/comm.cu); torch.distributed only carries the 128-byte NCCL id from rank 0 to the others when
the communicator is created.  The gloo path (CPU tests of the host logic) sums the float64 vectors.
"""
import ctypes

import numpy as np

from . import capi
from .capi import LAYOUT_SOA

BLOCK = 1024  # kReduceBlock (csrc/reduce.cuh)


def shard_range(N, rank, world):
    """[i0, i0+n): contiguous, whole reduction blocks, as even as possible."""
    blocks = (N + BLOCK - 1) // BLOCK
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    i0 = min(b0 * BLOCK, N)
    return i0, min(b1 * BLOCK, N) - i0


def tree_level1(part):
    """Level-1 tree over block sums (numpy restatement of ccu_tree_kernel; used on CPU tensors in the gloo tests)."""
    part = np.array(part, np.float64, copy=True)
    nb = part.shape[0]
    p2 = 1
    while p2 < nb:
        p2 *= 2
    pad = np.zeros((p2,) + part.shape[1:])
    pad[:nb] = part
    while pad.shape[0] > 1:
        pad = pad[0::2] + pad[1::2]
    return pad[0]


def combine_block_sums(part, group=None):
    """part: torch tensor (nblocks_global, nnz) holding this rank's block sums at their global rows, zeros
    elsewhere.  All-reduces in place (exact: disjoint supports) and returns it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


class ShardedCudaMap:
    """f.map(N, "cuda") over the GPUs of one node: this rank evaluates shard_range(N, rank, world)."""

    def __init__(self, tape, N, reduce_in=None, reduce_out=None, group=None):
        import torch
        import torch.distributed as dist
        self.f, self.N, self.group = tape, int(N), group
        init = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if init else 0
        self.world = dist.get_world_size(group) if init else 1
        self.i0, self.n = shard_range(self.N, self.rank, self.world)
        self.reduce_in = list(reduce_in) if reduce_in is not None else None
        self.reduce_out = list(reduce_out) if reduce_out is not None else None
        self.dev = torch.device("cuda", tape.device)
        nb = (self.N + BLOCK - 1) // BLOCK
        self.part = [torch.zeros((nb, nnz), dtype=torch.float64, device=self.dev)
                     if (self.reduce_out and self.reduce_out[j]) else None for j, nnz in enumerate(tape.nnz_out)]
        # the library's own NCCL communicator for this rank (ncclCommInitRank; id broadcast from rank 0)
        self.comm = None
        if self.world > 1 and any(p is not None for p in self.part):
            L = capi.lib()
            idbuf = (ctypes.c_ubyte * 128)()
            if self.rank == 0:
                capi.check(L.ccu_comm_unique_id(idbuf))
            backend = dist.get_backend(group)
            idt = torch.tensor(list(idbuf), dtype=torch.uint8, device=self.dev if backend == "nccl" else "cpu")
            dist.broadcast(idt, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            idbuf = (ctypes.c_ubyte * 128)(*idt.cpu().tolist())
            self.comm = L.ccu_comm_create_rank(idbuf, self.rank, self.world, tape.device)
            if not self.comm:
                raise capi.CcuError(capi.last_error())

    def close(self):
        if getattr(self, "comm", None):
            capi.lib().ccu_comm_destroy(self.comm)
            self.comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_device(self, d_arg, d_res, layout=LAYOUT_SOA, stream=None):
        """d_arg[j]: device address of this SHARD's input j (or of the single instance for reduce_in inputs);
        d_res[j]: shard output j, or -- for reduce_out outputs -- nnz_out[j] doubles receiving the global sum
        (identical on every rank)."""
        import torch
        L = capi.lib()
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(s):  # the zeroing is ordered with the block sums and the tree of the previous call
            for p in self.part:
                if p is not None:
                    p.zero_()
        ri, ro = capi.int_array(self.reduce_in), capi.int_array(self.reduce_out)
        parts = capi.ptr_array([None if p is None else p.data_ptr() for p in self.part])
        capi.check(L.ccu_map_eval_shard_device(
            self.f.handle, self.N, self.i0, self.n, capi.ptr_array(d_arg), capi.ptr_array(d_res),
            None if ri is None else ri.ctypes.data_as(capi.c_i_p), None if ro is None else ro.ctypes.data_as(capi.c_i_p),
            parts, layout, s.cuda_stream))
        for j, p in enumerate(self.part):
            if p is None:
                continue
            if self.comm:
                capi.check(L.ccu_comm_allreduce_block_sums(self.comm, capi.ptr_array([p.data_ptr()]), p.numel(),
                                                           capi.ptr_array([s.cuda_stream])))
            else:
                with torch.cuda.stream(s):
                    combine_block_sums(p, self.group)
            capi.check(L.ccu_reduce_tree_device(self.f.device, p.data_ptr(), self.N, self.f.nnz_out[j], d_res[j], s.cuda_stream))
