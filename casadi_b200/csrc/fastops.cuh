// Host-callable helper for the fast-path operators of ccu_ops.cuh (kernel in diag.cu).
#pragma once
#include <cuda_runtime.h>

namespace ccu {
// r[i] = the refined reciprocal the division fast path computes from the divisor c[i] alone (ccu::div_recip), evaluated
// ON THE CURRENT DEVICE so that a hoisted constant divisor gives exactly the bits of the in-line sequence
cudaError_t device_div_recip(const double* c, double* r, int n);
}  // namespace ccu
