// K5: fixed-shape sum over the instances of a mapped output.
//
// Replaces HorzRepsum::eval_gen (casadi/core/repmat.cpp:127-135) / MapSum::eval_gen
// (casadi/core/mapsum.cpp:170-184), which add the N blocks sequentially in index order.  A parallel
// sum cannot reproduce that rounding; instead the shape of the summation tree is FIXED by the
// instance index alone, so the result is bit-identical from run to run, for every launch geometry
// and for every number of GPUs:
//   level 0: instances are grouped in aligned blocks of kReduceBlock = 1024; inside a block the
//            values are combined by a balanced pairwise tree over the index bits (i, i^1), (i, i^2)..
//            (missing instances of the last block count as +0.0);
//   level 1: the block sums are combined by the same balanced pairwise tree over the block index,
//            padded with +0.0 to the next power of two.
// Multi-GPU: shards are whole blocks, every rank contributes its block sums at their global
// positions (zeros elsewhere, so an NCCL sum of the bit patterns is exact) and level 1 is evaluated on
// the merged vector -- see comm.cu and ccu_multi_eval_host (capi.cu).
#include "reduce.cuh"

namespace ccu {

// block sums: part[(blk - blk0)*nnz + k] = tree-sum_i x(i,k), i in block blk.
// x(i,k) at x[i*si + k*sk].  One CTA of 256 threads per (block, k-chunk).
__global__ void __launch_bounds__(256) ccu_block_sums_kernel(const double* __restrict__ x, long long si,
                                                             long long sk, long long N, int nnz,
                                                             double* __restrict__ part) {
  __shared__ double sh[kReduceBlock];
  const long long blk = blockIdx.x;
  const long long i0 = blk * kReduceBlock;
  for (int k = blockIdx.y; k < nnz; k += gridDim.y) {
    for (int t = threadIdx.x; t < kReduceBlock; t += blockDim.x) {
      long long i = i0 + t;
      sh[t] = i < N ? x[i * si + (long long)k * sk] : 0.0;
    }
    __syncthreads();
    // balanced pairwise tree: after step s, sh[t] (t multiple of 2s) holds the sum of 2s leaves
    for (int s = 1; s < kReduceBlock; s <<= 1) {
      for (int t = threadIdx.x; t < kReduceBlock / (2 * s); t += blockDim.x) {
        int lo = t * 2 * s;
        sh[lo] = sh[lo] + sh[lo + s];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) part[blk * nnz + k] = sh[0];
    __syncthreads();
  }
}

// level 1: out[k] = balanced pairwise tree over part[b*nnz + k], b in [0, nblocks), padded to pow2.
// Single CTA; works in place on a scratch copy `tmp` of nblocks_pow2*nnz doubles.
__global__ void __launch_bounds__(1024) ccu_tree_kernel(double* __restrict__ tmp, long long nblocks,
                                                        long long npow2, int nnz, double* __restrict__ out) {
  for (long long s = 1; s < npow2; s <<= 1) {
    const long long pairs = npow2 / (2 * s);
    for (long long idx = threadIdx.x; idx < pairs * nnz; idx += blockDim.x) {
      long long pr = idx / nnz;
      int k = (int)(idx % nnz);
      long long lo = pr * 2 * s, hi = lo + s;
      double a = lo < nblocks ? tmp[lo * nnz + k] : 0.0;
      double b = hi < nblocks ? tmp[hi * nnz + k] : 0.0;
      if (lo < nblocks) tmp[lo * nnz + k] = a + b;
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < nnz; k += blockDim.x) out[k] = nblocks > 0 ? tmp[k] : 0.0;
}

cudaError_t launch_block_sums(const double* x, long long si, long long sk, long long N, int nnz, double* part,
                              cudaStream_t stream) {
  if (N <= 0 || nnz <= 0) return cudaSuccess;
  long long nblocks = (N + kReduceBlock - 1) / kReduceBlock;
  dim3 grid((unsigned)nblocks, (unsigned)(nnz < 8 ? nnz : 8));
  ccu_block_sums_kernel<<<grid, 256, 0, stream>>>(x, si, sk, N, nnz, part);
  return cudaGetLastError();
}

cudaError_t launch_tree(double* tmp, long long nblocks, int nnz, double* out, cudaStream_t stream) {
  if (nnz <= 0) return cudaSuccess;
  long long npow2 = 1;
  while (npow2 < nblocks) npow2 <<= 1;
  ccu_tree_kernel<<<1, 1024, 0, stream>>>(tmp, nblocks, npow2, nnz, out);
  return cudaGetLastError();
}

}  // namespace ccu
