// K2: AoS <-> SoA layout conversion on the device (layout.cu).
#pragma once
#include <cuda_runtime.h>

namespace ccu {

// aos: n instances x nnz doubles, instance-major (the reference's Map layout, casadi/core/map.cpp:149-154)
// soa: nnz rows of `ld` doubles, element k of instance i at soa[k*ld + i]
cudaError_t launch_aos_to_soa(const double* aos, double* soa, long long n, int nnz, long long ld, cudaStream_t stream);
cudaError_t launch_soa_to_aos(const double* soa, double* aos, long long n, int nnz, long long ld, cudaStream_t stream);

}  // namespace ccu
