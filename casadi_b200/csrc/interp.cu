// K1: the SX tape interpreter for sm_100a.
//
// Replaces the `for (auto&& e : algorithm_) switch (e.op)` loop of SXFunction::eval
// (casadi/core/sx_function.cpp:111-124) inside the instance loop of Map::eval_gen
// (casadi/core/map.cpp:147-155).
//
// Execution model
//   * one instance per thread-lane, IPT lanes per thread (independent instances -> ILP);
//   * the program is PRE-DECODED on the host into 16-byte records (XInstr, interp.cuh): dispatch index, and
//     the byte offsets of the destination / source slots inside the thread's shared-memory column, so the
//     kernel does no field extraction and only 32-bit shared-memory address arithmetic;
//   * the records are staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double
//     buffered, one elected thread) and read with one warp-uniform LDS.128 per instruction; control flow never
//     diverges: SX tapes are straight-line, if_else is arithmetic;
//   * the work vector lives in shared memory as w[slot][lane]: lane-contiguous, hence bank-conflict-free
//     64-bit accesses; values the allocator could not keep in the shared slots are moved by FILL/SPILL to a
//     global scratch laid out [CTA][slot][lane] (coalesced);
//   * the previous result is forwarded in a register (X_ACC / X_NONE); the hot operations (+,-,*, neg, sq,
//     twice) have one switch case per operand-source combination, so their bodies are LDS/LDS/DADD/STS;
//   * persistent grid: each CTA loops over tiles of threads*IPT instances, so the scratch is sized by the
//     number of resident lanes, not by N.
// Compiled with -fmad=false (see ccu_ops.cuh for the rounding contract).
#include "interp.cuh"

#include "ccu_isa.h"
#include "ccu_ops.cuh"

namespace ccu {

#define CCU_FOR_E _Pragma("unroll") for (int e = 0; e < IPT; ++e)

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// The libm-class operations are called out of line, once per lane: inlining them IPT times into the dispatch loop
// bloats it, and looping over the lanes would index acc/x/y dynamically and push them to local memory.
__device__ __noinline__ double libm_unary(uint32_t op, double x) {
  switch (op) {
    case D_EXP: return exp(x);
    case D_LOG: return log(x);
    case D_SIN: return sin(x);
    case D_COS: return cos(x);
    case D_TAN: return tan(x);
    case D_ASIN: return asin(x);
    case D_ACOS: return acos(x);
    case D_ATAN: return atan(x);
    case D_ERF: return erf(x);
    case D_SINH: return sinh(x);
    case D_COSH: return cosh(x);
    case D_TANH: return tanh(x);
    case D_ASINH: return asinh(x);
    case D_ACOSH: return acosh(x);
    case D_ATANH: return atanh(x);
    case D_ERFINV: return op_erfinv(x);
    case D_LOG1P: return log1p(x);
    case D_EXPM1: return expm1(x);
    default: return CCU_NAN;
  }
}
__device__ __noinline__ double libm_binary(uint32_t op, double x, double y) {
  switch (op) {
    case D_POW: return pow(x, y);
    case D_FMOD: return fmod(x, y);
    case D_REMAINDER: return remainder(x, y);
    case D_ATAN2: return atan2(x, y);
    case D_HYPOT: return hypot(x, y);
    default: return CCU_NAN;
  }
}

}  // namespace

// dynamic shared memory: [ work vector: slots * WS doubles | 2 program stages of kChunk (+1 pad) records | 2 mbarriers ]
template <int IPT, bool SCRATCH>
__global__ void __launch_bounds__(IPT >= 4 ? 512 : 1024) ccu_interp_kernel(const XInstr* __restrict__ prog, const int nchunks, const IoDesc io,
                                                          const long long N, double* __restrict__ scratch,
                                                          const long long ntiles, const uint32_t stage_off,
                                                          const int slots_global) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int BD = blockDim.x;
  const int WS = IPT * BD;  // lanes per CTA
  const int tid = threadIdx.x;
  const uint32_t s_base = smem_u32(smem);
  uint32_t wcol = s_base + tid * 8u;  // this thread's column of the work vector; lane e at + e*BD*8
  asm volatile("" : "+r"(wcol));      // keep it in a register: the compiler would re-derive it from S2R every instruction
  const uint32_t lane_stride = static_cast<uint32_t>(BD) * 8u;
  const uint32_t stage0 = s_base + stage_off;
  const uint32_t bar0 = stage0 + 2u * kStageBytes;
  // blocked by CTA, [CTA][slot][lane]: a CTA's scratch is one contiguous region (the flat [slot][all lanes] layout put
  // consecutive slots megabytes apart: a TLB entry per slot)
  const long long scratch_stride = WS;
  double* const my_scratch = scratch + (long long)blockIdx.x * ((long long)slots_global * WS) + tid;

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t chunk_count = 0;  // chunks consumed so far by this CTA: stage = count & 1, parity = (count >> 1) & 1

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
    // first chunk of the program for this tile
    if (tid == 0) {
      const uint32_t st = chunk_count & 1u;
      mbar_expect_tx(bar0 + 8 * st, kChunk * sizeof(XInstr));
      tma_bulk_g2s(stage0 + st * kStageBytes, prog, kChunk * sizeof(XInstr), bar0 + 8 * st);
    }
    bool done = false;
    for (int c = 0; c < nchunks && !done; ++c, ++chunk_count) {
      const uint32_t st = chunk_count & 1u;
      // prefetch the next chunk into the other stage: every warp left it at the __syncthreads below
      if (tid == 0 && c + 1 < nchunks) {
        const uint32_t sn = st ^ 1u;
        mbar_expect_tx(bar0 + 8 * sn, kChunk * sizeof(XInstr));
        tma_bulk_g2s(stage0 + sn * kStageBytes, prog + (size_t)(c + 1) * kChunk, kChunk * sizeof(XInstr), bar0 + 8 * sn);
      }
      mbar_wait(bar0 + 8 * st, (chunk_count >> 1) & 1u);
      uint32_t pc = stage0 + st * kStageBytes;
      const uint32_t pc_end = pc + kChunk * sizeof(XInstr);
      uint4 ins = lds_u128(pc);
      while (pc < pc_end) {
        pc += sizeof(XInstr);
        const uint4 cur = ins;
        ins = lds_u128(pc);  // software prefetch of the next record (each stage has one padding record)
        const uint32_t fd = cur.y, fa = cur.z, fb = cur.w;
        double x[IPT], y[IPT], r[IPT];
        switch (cur.x & kXOpMask) {
          // ---- hot binary operations, one case per operand-source combination -------------------------------
#define CCU_BIN_CASES(NAME, EXPR)                                                                                   \
  case NAME##_MMS: CCU_FOR_E { x[e] = lds_f64(wcol + fa + e * lane_stride); y[e] = lds_f64(wcol + fb + e * lane_stride); } \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break;                   \
  case NAME##_MMN: CCU_FOR_E { x[e] = lds_f64(wcol + fa + e * lane_stride); y[e] = lds_f64(wcol + fb + e * lane_stride); } \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; } break;                                                               \
  case NAME##_AMS: CCU_FOR_E { x[e] = acc[e]; y[e] = lds_f64(wcol + fb + e * lane_stride); }                        \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break;                   \
  case NAME##_AMN: CCU_FOR_E { x[e] = acc[e]; y[e] = lds_f64(wcol + fb + e * lane_stride); }                        \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; } break;                                                               \
  case NAME##_MAS: CCU_FOR_E { x[e] = lds_f64(wcol + fa + e * lane_stride); y[e] = acc[e]; }                        \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break;                   \
  case NAME##_MAN: CCU_FOR_E { x[e] = lds_f64(wcol + fa + e * lane_stride); y[e] = acc[e]; }                        \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; } break;                                                               \
  case NAME##_AAS: CCU_FOR_E { x[e] = acc[e]; y[e] = acc[e]; }                                                      \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break;                   \
  case NAME##_AAN: CCU_FOR_E { x[e] = acc[e]; y[e] = acc[e]; }                                                      \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; } break;
          CCU_BIN_CASES(X_ADD, x[e] + y[e])
          CCU_BIN_CASES(X_SUB, x[e] - y[e])
          CCU_BIN_CASES(X_MUL, x[e] * y[e])
#undef CCU_BIN_CASES
#define CCU_UN_CASES(NAME, EXPR)                                                                         \
  case NAME##_MS: CCU_FOR_E x[e] = lds_f64(wcol + fa + e * lane_stride);                                  \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break;         \
  case NAME##_MN: CCU_FOR_E x[e] = lds_f64(wcol + fa + e * lane_stride);                                  \
    CCU_FOR_E { r[e] = EXPR; acc[e] = r[e]; } break;                                                     \
  case NAME##_AS: CCU_FOR_E { x[e] = acc[e]; r[e] = EXPR; acc[e] = r[e]; sts_f64(wcol + fd + e * lane_stride, r[e]); } break; \
  case NAME##_AN: CCU_FOR_E { x[e] = acc[e]; r[e] = EXPR; acc[e] = r[e]; } break;
          CCU_UN_CASES(X_NEG, -x[e])
          CCU_UN_CASES(X_SQ, x[e] * x[e])
          CCU_UN_CASES(X_TWICE, 2. * x[e])
#undef CCU_UN_CASES
          // ---- data movement ---------------------------------------------------------------------------------
          case X_CONST: {
            const double cst = __hiloint2double(static_cast<int>(fb), static_cast<int>(fa));
            CCU_FOR_E acc[e] = cst;
            if (fd != X_NONE) CCU_FOR_E sts_f64(wcol + fd + e * lane_stride, cst);
          } break;
          case X_INPUT: {
            const double* base = io.in[fa];
            if (base == nullptr) {  // NULL argument reads as zero (sx_function.cpp:116)
              CCU_FOR_E acc[e] = 0.0;
            } else {
              const long long si = io.in_si[fa];
              base += (long long)fb * io.in_sk[fa];
              CCU_FOR_E acc[e] = __ldg(base + inst[e] * si);
            }
            if (fd != X_NONE) CCU_FOR_E sts_f64(wcol + fd + e * lane_stride, acc[e]);
          } break;
          case X_OUTPUT: {
            double* base = io.out[fd];
            if (base != nullptr) {  // NULL result is not computed (sx_function.cpp:117)
              if (fa == X_ACC) { CCU_FOR_E x[e] = acc[e]; }
              else { CCU_FOR_E x[e] = lds_f64(wcol + fa + e * lane_stride); }
              const long long si = io.out_si[fd];
              base += (long long)fb * io.out_sk[fd];
              CCU_FOR_E { if (active[e]) base[inst[e] * si] = x[e]; }
            }
          } break;
          case X_FILL: {
            if (SCRATCH) {
              const double* g = my_scratch + (long long)fa * scratch_stride;
              CCU_FOR_E acc[e] = g[e * BD];
              if (fd != X_NONE) CCU_FOR_E sts_f64(wcol + fd + e * lane_stride, acc[e]);
            }
          } break;
          case X_SPILL: {
            if (SCRATCH) {
              if (fb == X_ACC) { CCU_FOR_E x[e] = acc[e]; }
              else { CCU_FOR_E x[e] = lds_f64(wcol + fb + e * lane_stride); }
              double* g = my_scratch + (long long)fa * scratch_stride;
              CCU_FOR_E g[e * BD] = x[e];
            }
          } break;
          case X_END: done = true; pc = pc_end; break;
          // ---- everything else: generic operand fetch, operation = the DevOp in the high half of `op` ----------
          case X_GENERIC_BIN: {
            const uint32_t op = cur.x >> 16;
            if (fa == X_ACC) { CCU_FOR_E x[e] = acc[e]; }
            else { CCU_FOR_E x[e] = lds_f64(wcol + fa + e * lane_stride); }
            if (fb == X_ACC) { CCU_FOR_E y[e] = acc[e]; }
            else { CCU_FOR_E y[e] = lds_f64(wcol + fb + e * lane_stride); }
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
              default: CCU_FOR_E r[e] = libm_binary(op, x[e], y[e]); break;  // one call per lane, static register indices
            }
            CCU_FOR_E acc[e] = r[e];
            if (fd != X_NONE) CCU_FOR_E sts_f64(wcol + fd + e * lane_stride, r[e]);
          } break;
          case X_GENERIC_UN: {
            const uint32_t op = cur.x >> 16;
            if (fa == X_ACC) { CCU_FOR_E x[e] = acc[e]; }
            else { CCU_FOR_E x[e] = lds_f64(wcol + fa + e * lane_stride); }
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
              default: CCU_FOR_E r[e] = libm_unary(op, x[e]); break;
            }
            CCU_FOR_E acc[e] = r[e];
            if (fd != X_NONE) CCU_FOR_E sts_f64(wcol + fd + e * lane_stride, r[e]);
          } break;
          default: __builtin_unreachable();
        }
      }
      __syncthreads();  // all warps are done with stage st: it may be refilled (chunk c+2) from the next iteration on
    }
  }
}

// host: translate the packed words of tape_compile (ccu_isa.h) into pre-decoded records for WS lanes per CTA
std::vector<XInstr> predecode(const std::vector<uint64_t>& words, int lanes_per_cta) {
  std::vector<XInstr> out;
  out.reserve(words.size() + kChunk);
  const uint32_t slot_bytes = static_cast<uint32_t>(lanes_per_cta) * 8u;
  auto off = [&](uint32_t f, uint32_t none) { return f == none ? X_NONE : f * slot_bytes; };
  for (size_t k = 0; k < words.size(); ++k) {
    const uint64_t w = words[k];
    const uint32_t op = CCU_DEC_OP(w), d = CCU_DEC_D(w), a = CCU_DEC_A(w), b = CCU_DEC_B(w);
    XInstr x{0, 0, 0, 0};
    const uint32_t fd = off(d, D_NONE), fa = off(a, F_ACC), fb = off(b, F_ACC);
    const bool st = d != D_NONE, am = a != F_ACC, bm = b != F_ACC;
    switch (op) {
      case D_END: x.op = X_END; break;
      case D_CONST: {
        const uint64_t bits = words[++k];
        x.op = X_CONST; x.d = fd; x.a = static_cast<uint32_t>(bits); x.b = static_cast<uint32_t>(bits >> 32);
      } break;
      case D_INPUT: x.op = X_INPUT; x.d = fd; x.a = a; x.b = b; break;
      case D_OUTPUT: x.op = X_OUTPUT; x.d = d; x.a = fa; x.b = b; break;
      case D_FILL: x.op = X_FILL; x.d = fd; x.a = a; break;
      case D_SPILL: x.op = X_SPILL; x.a = a; x.b = fb; break;
      case D_ADD: case D_SUB: case D_MUL: {
        // variant order inside a block of 8: MMS MMN AMS AMN MAS MAN AAS AAN
        const uint32_t base = op == D_ADD ? X_ADD_MMS : op == D_SUB ? X_SUB_MMS : X_MUL_MMS;
        x.op = base + (am ? 0u : 2u) + (bm ? 0u : 4u) + (st ? 0u : 1u);
        x.d = fd; x.a = fa; x.b = fb;
      } break;
      case D_NEG: case D_SQ: case D_TWICE: {
        // variant order inside a block of 4: MS MN AS AN
        const uint32_t base = op == D_NEG ? X_NEG_MS : op == D_SQ ? X_SQ_MS : X_TWICE_MS;
        x.op = base + (am ? 0u : 2u) + (st ? 0u : 1u);
        x.d = fd; x.a = fa; x.b = fa;
      } break;
      default: x.op = (op < D_UN_FIRST ? X_GENERIC_BIN : X_GENERIC_UN) | (op << 16); x.d = fd; x.a = fa; x.b = fb; break;
    }
    out.push_back(x);
  }
  while (out.empty() || out.size() % kChunk != 0) out.push_back(XInstr{X_END, 0, 0, 0});
  return out;
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

size_t plan_smem_bytes(const LaunchPlan& plan) {
  size_t work = (size_t)plan.slots_shared * plan.ipt * plan.threads * sizeof(double);
  work = (work + 15) / 16 * 16;
  return work + 2 * kStageBytes + 16;
}

cudaError_t plan_occupancy(LaunchPlan* plan, int device) {
  plan->smem_bytes = plan_smem_bytes(*plan);
  const bool sc = plan->slots_global > 0;
  switch (plan->ipt) {
    case 1: return sc ? occupancy_for<1, true>(plan, device) : occupancy_for<1, false>(plan, device);
    case 2: return sc ? occupancy_for<2, true>(plan, device) : occupancy_for<2, false>(plan, device);
    case 4: return sc ? occupancy_for<4, true>(plan, device) : occupancy_for<4, false>(plan, device);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_interp(const LaunchPlan& plan, const XInstr* d_prog, long long n_records, const IoDesc& io, long long N,
                          double* d_scratch, cudaStream_t stream) {
  if (N <= 0) return cudaSuccess;
  const long long lanes = (long long)plan.threads * plan.ipt;
  const long long ntiles = (N + lanes - 1) / lanes;
  const int grid = (int)(ntiles < plan.grid ? ntiles : plan.grid);
  const bool sc = plan.slots_global > 0;
  const int nchunks = static_cast<int>(n_records / kChunk);
  const uint32_t stage_off = static_cast<uint32_t>(plan.smem_bytes - 2 * kStageBytes - 16);
#define CCU_LAUNCH(I, S)                                                                                              \
  ccu_interp_kernel<I, S><<<grid, plan.threads, plan.smem_bytes, stream>>>(d_prog, nchunks, io, N, d_scratch, ntiles, \
                                                                           stage_off, plan.slots_global)
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
