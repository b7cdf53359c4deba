// FP64 issue-rate microbenchmark: independent DADD / DMUL / DFMA chains per thread, no memory traffic.
// Gives the denominators for the FP64 side of the roofline (non-FMA ops/s is what a -fmad=false kernel can reach).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a); x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a); }
    if (MODE == 1) { x0 = __dmul_rn(x0, b); x1 = __dmul_rn(x1, b); x2 = __dmul_rn(x2, b); x3 = __dmul_rn(x3, b); x4 = __dmul_rn(x4, b); x5 = __dmul_rn(x5, b); x6 = __dmul_rn(x6, b); x7 = __dmul_rn(x7, b); }
    if (MODE == 2) { x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a); x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * 8, block = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* nm[3] = {"dadd", "dmul", "dfma"};
  for (int m = 0; m < 3; ++m) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<grid, block>>>(out, iters, 1e-9, 1.0000001);
      if (m == 1) k<1><<<grid, block>>>(out, iters, 1e-9, 1.0000001);
      if (m == 2) k<2><<<grid, block>>>(out, iters, 1e-9, 1.0000001);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)grid * block * iters * 8;
    printf("{\"op\": \"%s\", \"sms\": %d, \"ms\": %.3f, \"Gop_s\": %.1f}\n", nm[m], sms, best, ops / best / 1e6);
  }
  return 0;
}
