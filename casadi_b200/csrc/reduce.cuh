// Launch interface of the fixed-shape reduction kernels (reduce.cu).
#pragma once
#include <cuda_runtime.h>

namespace ccu {

constexpr int kReduceBlock = 1024;  // instances per level-0 block; shard boundaries must be multiples of it

// part[b*nnz + k] = pairwise-tree sum over block b of x(i,k) = x[i*si + k*sk]
cudaError_t launch_block_sums(const double* x, long long si, long long sk, long long N, int nnz, double* part,
                              cudaStream_t stream);
// out[k] = pairwise tree over the nblocks block sums in tmp (destroyed)
cudaError_t launch_tree(double* tmp, long long nblocks, int nnz, double* out, cudaStream_t stream);

}  // namespace ccu
