// K2: AoS <-> SoA conversion.  The reference's Map hands every instance a contiguous slice
// arg[j] + i*nnz_in(j) (casadi/core/map.cpp:149-154); one-instance-per-thread kernels want element k of
// consecutive instances adjacent.  A 32x32 tile goes through shared memory (padded: no bank conflicts), so
// both the global reads and the global writes are coalesced.  Pure data movement: HBM-bound, 16 B per value.
#include "layout.cuh"

namespace ccu {
namespace {

constexpr int kTile = 32;
constexpr int kRows = 8;

// TO_SOA:  in = aos[i*nnz + k]  -> out = soa[k*ld + i]
// !TO_SOA: in = soa[k*ld + i]   -> out = aos[i*nnz + k]
template <bool TO_SOA>
__global__ void __launch_bounds__(kTile* kRows) ccu_layout_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                  long long n, int nnz, long long ld) {
  __shared__ double tile[kTile][kTile + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = blockIdx.y * kTile;
  for (long long i0 = (long long)blockIdx.x * kTile; i0 < n; i0 += (long long)gridDim.x * kTile) {
    if (TO_SOA) {
#pragma unroll
      for (int r = 0; r < kTile; r += kRows) {  // rows = instances, columns = k (contiguous in AoS)
        const long long i = i0 + ty + r;
        const int k = k0 + tx;
        if (i < n && k < nnz) tile[ty + r][tx] = in[i * nnz + k];
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kTile; r += kRows) {  // rows = k, columns = instances (contiguous in SoA)
        const int k = k0 + ty + r;
        const long long i = i0 + tx;
        if (i < n && k < nnz) out[(long long)k * ld + i] = tile[tx][ty + r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < kTile; r += kRows) {
        const int k = k0 + ty + r;
        const long long i = i0 + tx;
        if (i < n && k < nnz) tile[ty + r][tx] = in[(long long)k * ld + i];
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kTile; r += kRows) {
        const long long i = i0 + ty + r;
        const int k = k0 + tx;
        if (i < n && k < nnz) out[i * nnz + k] = tile[tx][ty + r];
      }
    }
    __syncthreads();
  }
}

template <bool TO_SOA>
cudaError_t launch(const double* in, double* out, long long n, int nnz, long long ld, cudaStream_t stream) {
  if (n <= 0 || nnz <= 0) return cudaSuccess;
  long long gx = (n + kTile - 1) / kTile;
  if (gx > 148 * 64) gx = 148 * 64;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>((nnz + kTile - 1) / kTile));
  ccu_layout_kernel<TO_SOA><<<grid, dim3(kTile, kRows), 0, stream>>>(in, out, n, nnz, ld);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_aos_to_soa(const double* aos, double* soa, long long n, int nnz, long long ld, cudaStream_t stream) {
  return launch<true>(aos, soa, n, nnz, ld, stream);
}
cudaError_t launch_soa_to_aos(const double* soa, double* aos, long long n, int nnz, long long ld, cudaStream_t stream) {
  return launch<false>(soa, aos, n, nnz, ld, stream);
}

}  // namespace ccu
