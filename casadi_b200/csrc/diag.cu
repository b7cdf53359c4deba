// Diagnostics: the FP64 side of the roofline, measured on the device the library runs on.
//
// T_fp64 = flops_per_eval / P64 with P64 the FP64 NON-FMA issue rate (the library is compiled
// with -fmad=false, so one tape instruction is at best one DADD/DMUL).  MEASURED_PEAKS.json carries HBM and
// bf16 numbers only, so P64 is measured here: independent DADD chains, no memory traffic, CUDA-event timed.
#include "../../include/casadi_cuda.h"

#include <cuda_runtime.h>

namespace {

__global__ void __launch_bounds__(256) ccu_dadd_rate_kernel(double* out, int iters, double a) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a);
    x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

extern "C" int ccu_fp64_issue_rate(int device, double* ops_per_s) {
  if (!ops_per_s) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 1;
  const int grid = sms * 8, block = 256, iters = 20000;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * grid * block) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    ccu_dadd_rate_kernel<<<grid, block>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return 1;
  *ops_per_s = static_cast<double>(grid) * block * iters * 8 / (best * 1e-3);
  return 0;
}
