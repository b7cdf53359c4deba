// Launch interface of the tape-interpreter kernel (interp.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ccu {

constexpr int kMaxIO = 32;  // inputs / outputs passed by kernel parameter

// element k of instance i of input j lives at in[j][i*in_si[j] + k*in_sk[j]]
//   AoS (reference layout, map.cpp:149-154): si = nnz, sk = 1
//   SoA (device fast path):                  si = 1,   sk = N
//   reduce_in broadcast (repmat.cpp:44-50):  si = 0,   sk = 1
struct IoDesc {
  const double* in[kMaxIO];
  long long in_si[kMaxIO];
  long long in_sk[kMaxIO];
  double* out[kMaxIO];
  long long out_si[kMaxIO];
  long long out_sk[kMaxIO];
};

struct LaunchPlan {
  int threads = 128;     // CTA size
  int ipt = 1;           // instances per thread (1, 2 or 4)
  int slots_shared = 0;  // shared work slots per instance
  int slots_global = 0;  // scratch slots per instance
  int ctas_per_sm = 0;   // filled by plan_occupancy
  int grid = 0;          // persistent grid size
  size_t smem_bytes = 0;
};

// fills ctas_per_sm/grid/smem_bytes for the current device; returns cudaSuccess or an error
cudaError_t plan_occupancy(LaunchPlan* plan, int device);

// scratch must hold slots_global * grid * threads * ipt doubles
cudaError_t launch_interp(const LaunchPlan& plan, const uint64_t* d_prog, const IoDesc& io, long long N,
                          double* d_scratch, cudaStream_t stream);

}  // namespace ccu
