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

// ------------------------------------------------------------------------------------------------------------
// Fast-path operators of the specialised kernels (ccu_ops.cuh): the divisor-only half of the division on the
// device (constant divisors are hoisted by the code generator) and a self test against the plain operators.
#include "ccu_ops.cuh"
#include "fastops.cuh"

namespace {

__global__ void ccu_div_recip_kernel(const double* c, double* r, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = ccu::div_recip(c[i]);
}

__device__ __forceinline__ unsigned long long ccu_mix(unsigned long long z) {  // splitmix64
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// operand classes: 0 = raw bit patterns (every exponent, nan, inf, denormals), 1 = moderate magnitudes (the regime of
// the benchmark tapes), 2 = trig arguments up to 2^33 (crosses the fast-path limit), 3 = special values
__device__ double ccu_operand(unsigned long long h, int cls) {
  if (cls == 0) return __longlong_as_double(static_cast<long long>(h));
  if (cls == 1) {
    const double m = 1.0 + static_cast<double>(h >> 12) * 2.220446049250313e-16;
    const int e = static_cast<int>((h >> 4) & 63) - 32;
    return ((h & 1) ? -m : m) * exp2(static_cast<double>(e));
  }
  if (cls == 2) {
    const double m = static_cast<double>(h >> 11) * 1.1102230246251565e-16;
    return ((h & 1) ? -m : m) * exp2(static_cast<double>((h >> 1) % 34));
  }
  const double sp[12] = {0.0, -0.0, 1.0, -1.0, CCU_INF, -CCU_INF, CCU_NAN, 4.9e-324, 2.2250738585072014e-308,
                         1.7976931348623157e308, 2147483648.0, 3.141592653589793};
  return sp[h % 12];
}

__global__ void ccu_fastops_selftest_kernel(long long n, unsigned long long seed, unsigned long long* counts) {
  unsigned long long mism = 0, flagged = 0, checked = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long h0 = ccu_mix(seed + 2 * i), h1 = ccu_mix(seed + 2 * i + 1);
    const int cls = static_cast<int>(i & 3);
    const double a = ccu_operand(h0, cls == 2 ? 1 : cls), b = ccu_operand(h1, cls == 2 ? 1 : cls), x = ccu_operand(h0, cls == 1 ? 2 : cls);
    auto same = [](double u, double v) { return __double_as_longlong(u) == __double_as_longlong(v) || (u != u && v != v); };
    {  // division, general and with the divisor half hoisted
      bool bad = false;
      const double q = ccu::div_fast(a, b, bad);
      bool bad2 = false;
      const double q2 = ccu::div_finish(a, b, ccu::div_recip(b), bad2);
      ++checked;
      if (bad) ++flagged;
      else if (!same(q, a / b) || !same(q2, q) || bad2) ++mism;
    }
    {  // sincos / sin / cos
      bool bad = false;
      double s, c;
      ccu::sincos_fast(x, &s, &c, bad);
      const double s1 = ccu::trig_one_fast(x, 0, bad), c1 = ccu::trig_one_fast(x, 1, bad);
      double sr, cr;
      sincos(x, &sr, &cr);
      ++checked;
      if (bad) ++flagged;
      else if (!same(s, sr) || !same(c, cr) || !same(s1, sin(x)) || !same(c1, cos(x))) ++mism;
    }
  }
  atomicAdd(&counts[0], mism);
  atomicAdd(&counts[1], flagged);
  atomicAdd(&counts[2], checked);
}

}  // namespace

namespace ccu {
cudaError_t device_div_recip(const double* c, double* r, int n) {
  if (n <= 0) return cudaSuccess;
  double *dc = nullptr, *dr = nullptr;
  cudaError_t e = cudaMalloc(&dc, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&dr, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMemcpy(dc, c, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    ccu_div_recip_kernel<<<(n + 127) / 128, 128>>>(dc, dr, n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(r, dr, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (dc) cudaFree(dc);
  if (dr) cudaFree(dr);
  return e;
}
}  // namespace ccu

extern "C" int ccu_selftest_fastops(int device, long long n, unsigned long long seed, unsigned long long counts[3]) {
  if (!counts || n < 0) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 3 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  cudaMemset(d, 0, 3 * sizeof(unsigned long long));
  ccu_fastops_selftest_kernel<<<148 * 8, 256>>>(n, seed, d);
  const cudaError_t e = cudaMemcpy(counts, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : 1;
}
