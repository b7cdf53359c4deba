// FP64 latency / ILP microbenchmark (design aid): how many independent dependent-chains per warp and warps per
// scheduler does a non-FMA FP64 stream need to saturate the pipe, and what do division / sincos cost?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/fp64_ilp.cu -o tools/fp64_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = __dmul_rn(x[j], b);
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = __dadd_rn(x[j], a);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: x = a / x (general division), 1: x = x / c (runtime-uniform c: general division), 2: sincos, 3: sin + cos separately
// 4: reciprocal-multiply + one correction (3 FP64 instructions) for a constant divisor
template <int MODE, int ILP>
__global__ void k_fun(double* out, int iters, double a, double c, double rc) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = 1.0 + 1e-3 * (threadIdx.x + j);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (MODE == 0) x[j] = a / x[j];
      if (MODE == 1) x[j] = x[j] / c;
      if (MODE == 2) { double s, co; sincos(x[j], &s, &co); x[j] = s + co; }
      if (MODE == 3) { x[j] = sin(x[j]) + cos(x[j]); }
      if (MODE == 4) { double q = __dmul_rn(x[j], rc); double r = __fma_rn(-q, c, x[j]); x[j] = __fma_rn(r, rc, q); }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 2);
  const int iters = 4000;
  const int warps[] = {4, 8, 16, 32};
  for (int w : warps) {
    const int block = 32 * w;
#define RUN(I) { float ms = time_it([&] { k_chain<I><<<sms, block>>>(out, iters, 1e-9, 1.0000001); }); \
      double ops = (double)sms * block * iters * 8.0 * I; \
      printf("{\"bench\": \"chain\", \"warps_per_sm\": %d, \"ilp\": %d, \"Top_s\": %.2f}\n", w, I, ops / ms / 1e9); }
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(6) RUN(8)
#undef RUN
  }
  const char* names[] = {"div a/x", "div x/c", "sincos", "sin+cos", "x/c as mul+2fma"};
  for (int w : {8, 16, 32}) {
    const int block = 32 * w;
    const int it = 400;
    float ms[5];
    ms[0] = time_it([&] { k_fun<0, 4><<<sms, block>>>(out, it, 1.7, 1.2, 1 / 1.2); });
    ms[1] = time_it([&] { k_fun<1, 4><<<sms, block>>>(out, it, 1.7, 1.0000001, 1 / 1.0000001); });
    ms[2] = time_it([&] { k_fun<2, 4><<<sms, block>>>(out, it, 1.7, 1.2, 1 / 1.2); });
    ms[3] = time_it([&] { k_fun<3, 4><<<sms, block>>>(out, it, 1.7, 1.2, 1 / 1.2); });
    ms[4] = time_it([&] { k_fun<4, 4><<<sms, block>>>(out, it, 1.7, 1.0000001, 1 / 1.0000001); });
    for (int m = 0; m < 5; ++m) {
      double calls = (double)sms * block * it * 4;
      // cost of one call in units of one FP64 issue slot at the measured 18.4 T/s peak
      printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"Gcalls_s\": %.2f, \"fp64_slots_per_call\": %.1f}\n", names[m], w,
             calls / ms[m] / 1e6, 18.4e12 / (calls / (ms[m] * 1e-3)));
    }
  }
  return 0;
}
