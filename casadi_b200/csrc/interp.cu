// K1: the SX tape interpreter for sm_100a.
//
// Replaces the `for (auto&& e : algorithm_) switch (e.op)` loop of SXFunction::eval
// (casadi/core/sx_function.cpp:111-124) inside the instance loop of Map::eval_gen
// (casadi/core/map.cpp:147-155).
//
// Execution model
//   * one instance per thread-lane, IPT lanes per thread (independent instances -> ILP);
//   * the program (ccu_isa.h) is read warp-uniformly (every thread decodes the same word, so the
//     control flow never diverges: SX tapes are straight-line, if_else is arithmetic);
//   * the work vector lives in shared memory as w[slot][lane]: lane-contiguous, hence
//     bank-conflict-free 64-bit accesses;  values the allocator could not keep in the shared
//     slots are moved by FILL/SPILL to a global scratch laid out [slot][resident lane] (coalesced);
//   * the previous result is forwarded in a register (F_ACC / D_NONE);
//   * persistent grid: each CTA loops over tiles of threads*IPT instances, so the scratch is sized
//     by the number of resident lanes, not by N.
// Compiled with -fmad=false (see ccu_ops.cuh for the rounding contract).
#include "interp.cuh"

#include "ccu_isa.h"
#include "ccu_ops.cuh"

namespace ccu {

#define CCU_FOR_E _Pragma("unroll") for (int e = 0; e < IPT; ++e)
#define CCU_FOR_E_ROLLED _Pragma("unroll 1") for (int e = 0; e < IPT; ++e)

template <int IPT, bool SCRATCH>
__global__ void __launch_bounds__(1024) ccu_interp_kernel(const uint64_t* __restrict__ prog, const IoDesc io,
                                                          const long long N, double* __restrict__ scratch,
                                                          const long long ntiles) {
  extern __shared__ double w[];  // [slot][IPT*blockDim.x]
  const int BD = blockDim.x;
  const int WS = IPT * BD;       // lanes per CTA
  const int tid = threadIdx.x;
  const long long scratch_stride = (long long)gridDim.x * WS;
  double* const my_scratch = scratch + (long long)blockIdx.x * WS + tid;
  double* const my_w = w + tid;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long inst[IPT];  // instance index of each lane, clamped so that inactive lanes stay in range
    bool active[IPT];
    CCU_FOR_E {
      long long gi = tile * WS + (long long)e * BD + tid;
      active[e] = gi < N;
      inst[e] = active[e] ? gi : N - 1;
    }
    double acc[IPT];
    CCU_FOR_E acc[e] = 0.0;
    const uint64_t* pc = prog;
    uint64_t word = __ldg(pc);
    for (;;) {
      const uint64_t next = __ldg(pc + 1);  // program is padded: pc+1 is always readable
      const uint32_t op = CCU_DEC_OP(word);
      const uint32_t fd = CCU_DEC_D(word);
      const uint32_t fa = CCU_DEC_A(word);
      const uint32_t fb = CCU_DEC_B(word);
      ++pc;
      word = next;
      if (op >= D_UN_FIRST) {
        // ---------------------------------------------------------------- unary
        double x[IPT], r[IPT];
        if (fa == F_ACC) { CCU_FOR_E x[e] = acc[e]; }
        else { const double* p = my_w + (size_t)fa * WS; CCU_FOR_E x[e] = p[e * BD]; }
        switch (op) {
          case D_COPY: CCU_FOR_E r[e] = x[e]; break;
          case D_NEG: CCU_FOR_E r[e] = -x[e]; break;
          case D_SQRT: CCU_FOR_E r[e] = sqrt(x[e]); break;
          case D_SQ: CCU_FOR_E r[e] = x[e] * x[e]; break;
          case D_TWICE: CCU_FOR_E r[e] = 2. * x[e]; break;
          case D_INV: CCU_FOR_E r[e] = 1. / x[e]; break;
          case D_FABS: CCU_FOR_E r[e] = fabs(x[e]); break;
          case D_SIGN: CCU_FOR_E r[e] = op_sign(x[e]); break;
          case D_NOT: CCU_FOR_E r[e] = op_not(x[e]); break;
          case D_FLOOR: CCU_FOR_E r[e] = floor(x[e]); break;
          case D_CEIL: CCU_FOR_E r[e] = ceil(x[e]); break;
          case D_EXP: CCU_FOR_E_ROLLED r[e] = exp(x[e]); break;
          case D_LOG: CCU_FOR_E_ROLLED r[e] = log(x[e]); break;
          case D_SIN: CCU_FOR_E_ROLLED r[e] = sin(x[e]); break;
          case D_COS: CCU_FOR_E_ROLLED r[e] = cos(x[e]); break;
          case D_TAN: CCU_FOR_E_ROLLED r[e] = tan(x[e]); break;
          case D_ASIN: CCU_FOR_E_ROLLED r[e] = asin(x[e]); break;
          case D_ACOS: CCU_FOR_E_ROLLED r[e] = acos(x[e]); break;
          case D_ATAN: CCU_FOR_E_ROLLED r[e] = atan(x[e]); break;
          case D_ERF: CCU_FOR_E_ROLLED r[e] = erf(x[e]); break;
          case D_SINH: CCU_FOR_E_ROLLED r[e] = sinh(x[e]); break;
          case D_COSH: CCU_FOR_E_ROLLED r[e] = cosh(x[e]); break;
          case D_TANH: CCU_FOR_E_ROLLED r[e] = tanh(x[e]); break;
          case D_ASINH: CCU_FOR_E_ROLLED r[e] = asinh(x[e]); break;
          case D_ACOSH: CCU_FOR_E_ROLLED r[e] = acosh(x[e]); break;
          case D_ATANH: CCU_FOR_E_ROLLED r[e] = atanh(x[e]); break;
          case D_ERFINV: CCU_FOR_E_ROLLED r[e] = op_erfinv(x[e]); break;
          case D_LOG1P: CCU_FOR_E_ROLLED r[e] = log1p(x[e]); break;
          case D_EXPM1: CCU_FOR_E_ROLLED r[e] = expm1(x[e]); break;
          default: CCU_FOR_E r[e] = CCU_NAN; break;
        }
        CCU_FOR_E acc[e] = r[e];
        if (fd != D_NONE) { double* p = my_w + (size_t)fd * WS; CCU_FOR_E p[e * BD] = r[e]; }
      } else if (op >= D_BIN_FIRST) {
        // ---------------------------------------------------------------- binary
        double x[IPT], y[IPT], r[IPT];
        if (fa == F_ACC) { CCU_FOR_E x[e] = acc[e]; }
        else { const double* p = my_w + (size_t)fa * WS; CCU_FOR_E x[e] = p[e * BD]; }
        if (fb == F_ACC) { CCU_FOR_E y[e] = acc[e]; }
        else { const double* p = my_w + (size_t)fb * WS; CCU_FOR_E y[e] = p[e * BD]; }
        switch (op) {
          case D_ADD: CCU_FOR_E r[e] = x[e] + y[e]; break;
          case D_SUB: CCU_FOR_E r[e] = x[e] - y[e]; break;
          case D_MUL: CCU_FOR_E r[e] = x[e] * y[e]; break;
          case D_DIV: CCU_FOR_E r[e] = x[e] / y[e]; break;
          case D_LT: CCU_FOR_E r[e] = x[e] < y[e] ? 1.0 : 0.0; break;
          case D_LE: CCU_FOR_E r[e] = x[e] <= y[e] ? 1.0 : 0.0; break;
          case D_EQ: CCU_FOR_E r[e] = x[e] == y[e] ? 1.0 : 0.0; break;
          case D_NE: CCU_FOR_E r[e] = x[e] != y[e] ? 1.0 : 0.0; break;
          case D_AND: CCU_FOR_E r[e] = op_and(x[e], y[e]); break;
          case D_OR: CCU_FOR_E r[e] = op_or(x[e], y[e]); break;
          case D_IF_ELSE_ZERO: CCU_FOR_E r[e] = op_if_else_zero(x[e], y[e]); break;
          case D_FMIN: CCU_FOR_E r[e] = op_fmin(x[e], y[e]); break;
          case D_FMAX: CCU_FOR_E r[e] = op_fmax(x[e], y[e]); break;
          case D_COPYSIGN: CCU_FOR_E r[e] = copysign(x[e], y[e]); break;
          case D_POW: CCU_FOR_E_ROLLED r[e] = pow(x[e], y[e]); break;
          case D_FMOD: CCU_FOR_E_ROLLED r[e] = fmod(x[e], y[e]); break;
          case D_REMAINDER: CCU_FOR_E_ROLLED r[e] = remainder(x[e], y[e]); break;
          case D_ATAN2: CCU_FOR_E_ROLLED r[e] = atan2(x[e], y[e]); break;
          case D_HYPOT: CCU_FOR_E_ROLLED r[e] = hypot(x[e], y[e]); break;
          default: CCU_FOR_E r[e] = CCU_NAN; break;
        }
        CCU_FOR_E acc[e] = r[e];
        if (fd != D_NONE) { double* p = my_w + (size_t)fd * WS; CCU_FOR_E p[e * BD] = r[e]; }
      } else if (op == D_CONST) {
        const double c = __longlong_as_double((long long)word);  // literal = the word after the CONST
        ++pc;
        word = __ldg(pc);
        CCU_FOR_E acc[e] = c;
        if (fd != D_NONE) { double* p = my_w + (size_t)fd * WS; CCU_FOR_E p[e * BD] = c; }
      } else if (op == D_INPUT) {
        const double* base = io.in[fa];
        if (base == nullptr) {  // NULL argument reads as zero (sx_function.cpp:116)
          CCU_FOR_E acc[e] = 0.0;
        } else {
          const long long si = io.in_si[fa];
          base += (long long)fb * io.in_sk[fa];
          CCU_FOR_E acc[e] = __ldg(base + inst[e] * si);
        }
        if (fd != D_NONE) { double* p = my_w + (size_t)fd * WS; CCU_FOR_E p[e * BD] = acc[e]; }
      } else if (op == D_OUTPUT) {
        double* base = io.out[fd];
        if (base != nullptr) {  // NULL result is not computed (sx_function.cpp:117)
          double x[IPT];
          if (fa == F_ACC) { CCU_FOR_E x[e] = acc[e]; }
          else { const double* p = my_w + (size_t)fa * WS; CCU_FOR_E x[e] = p[e * BD]; }
          const long long si = io.out_si[fd];
          base += (long long)fb * io.out_sk[fd];
          CCU_FOR_E { if (active[e]) base[inst[e] * si] = x[e]; }
        }
      } else if (SCRATCH && op == D_FILL) {
        const double* g = my_scratch + (long long)fa * scratch_stride;
        CCU_FOR_E acc[e] = g[e * BD];
        if (fd != D_NONE) { double* p = my_w + (size_t)fd * WS; CCU_FOR_E p[e * BD] = acc[e]; }
      } else if (SCRATCH && op == D_SPILL) {
        double x[IPT];
        if (fb == F_ACC) { CCU_FOR_E x[e] = acc[e]; }
        else { const double* p = my_w + (size_t)fb * WS; CCU_FOR_E x[e] = p[e * BD]; }
        double* g = my_scratch + (long long)fa * scratch_stride;
        CCU_FOR_E g[e * BD] = x[e];
      } else {
        break;  // D_END
      }
    }
  }
}

template <int IPT, bool SCRATCH>
static cudaError_t occupancy_for(LaunchPlan* plan, int device) {
  auto kern = ccu_interp_kernel<IPT, SCRATCH>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
  if (e != cudaSuccess) return e;
  int nb = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, plan->threads, plan->smem_bytes);
  if (e != cudaSuccess) return e;
  if (nb < 1) return cudaErrorInvalidConfiguration;
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return e;
  plan->ctas_per_sm = nb;
  plan->grid = nb * sms;
  return cudaSuccess;
}

cudaError_t plan_occupancy(LaunchPlan* plan, int device) {
  plan->smem_bytes = (size_t)plan->slots_shared * plan->ipt * plan->threads * sizeof(double);
  const bool sc = plan->slots_global > 0;
  switch (plan->ipt) {
    case 1: return sc ? occupancy_for<1, true>(plan, device) : occupancy_for<1, false>(plan, device);
    case 2: return sc ? occupancy_for<2, true>(plan, device) : occupancy_for<2, false>(plan, device);
    case 4: return sc ? occupancy_for<4, true>(plan, device) : occupancy_for<4, false>(plan, device);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_interp(const LaunchPlan& plan, const uint64_t* d_prog, const IoDesc& io, long long N,
                          double* d_scratch, cudaStream_t stream) {
  if (N <= 0) return cudaSuccess;
  const long long lanes = (long long)plan.threads * plan.ipt;
  const long long ntiles = (N + lanes - 1) / lanes;
  const int grid = (int)(ntiles < plan.grid ? ntiles : plan.grid);
  const bool sc = plan.slots_global > 0;
#define CCU_LAUNCH(I, S) \
  ccu_interp_kernel<I, S><<<grid, plan.threads, plan.smem_bytes, stream>>>(d_prog, io, N, d_scratch, ntiles)
  switch (plan.ipt) {
    case 1: if (sc) CCU_LAUNCH(1, true); else CCU_LAUNCH(1, false); break;
    case 2: if (sc) CCU_LAUNCH(2, true); else CCU_LAUNCH(2, false); break;
    case 4: if (sc) CCU_LAUNCH(4, true); else CCU_LAUNCH(4, false); break;
    default: return cudaErrorInvalidValue;
  }
#undef CCU_LAUNCH
  return cudaGetLastError();
}

}  // namespace ccu
