// Launch interface of the tape-interpreter kernel (interp.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

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

// Pre-decoded instruction record the kernel executes (built on the host by predecode() from the packed
// words of ccu_isa.h once the launch geometry is known).  d/a/b are BYTE offsets of the slot inside a
// thread's shared-memory column (slot * lanes_per_cta * 8), X_ACC / X_NONE for the forwarding register, or
// the raw index for INPUT / OUTPUT / FILL / SPILL; CONST carries the literal in (a = low, b = high word).
struct XInstr {
  uint32_t op, d, a, b;
};
static_assert(sizeof(XInstr) == 16, "XInstr is one LDS.128");

constexpr uint32_t X_ACC = 0xffffffffu;   // source: the previous result
constexpr uint32_t X_NONE = 0xffffffffu;  // destination: not stored
constexpr int kChunk = 128;               // records per TMA bulk copy (2 KB)
constexpr uint32_t kStageBytes = (kChunk + 1) * sizeof(XInstr);  // + one never-executed padding record

// dispatch indices: blocks of operand-source variants for the hot operations
//   binary  M = operand in shared memory, A = forwarding register; S = store result, N = register only
enum XOp : uint32_t {
  X_END = 0, X_CONST, X_INPUT, X_OUTPUT, X_FILL, X_SPILL,
  X_ADD_MMS = 6, X_ADD_MMN, X_ADD_AMS, X_ADD_AMN, X_ADD_MAS, X_ADD_MAN, X_ADD_AAS, X_ADD_AAN,
  X_SUB_MMS = 14, X_SUB_MMN, X_SUB_AMS, X_SUB_AMN, X_SUB_MAS, X_SUB_MAN, X_SUB_AAS, X_SUB_AAN,
  X_MUL_MMS = 22, X_MUL_MMN, X_MUL_AMS, X_MUL_AMN, X_MUL_MAS, X_MUL_MAN, X_MUL_AAS, X_MUL_AAN,
  X_NEG_MS = 30, X_NEG_MN, X_NEG_AS, X_NEG_AN,
  X_SQ_MS = 34, X_SQ_MN, X_SQ_AS, X_SQ_AN,
  X_TWICE_MS = 38, X_TWICE_MN, X_TWICE_AS, X_TWICE_AN,
  X_GENERIC_BIN = 42,  // generic operand fetch, two operands; the DevOp travels in the high half of `op`
  X_GENERIC_UN = 43,   // generic operand fetch, one operand
  X_COUNT = 44,
};
constexpr uint32_t kXOpMask = 0xffffu;  // low half of XInstr::op = dispatch index (dense 0..X_COUNT-1)

struct LaunchPlan {
  int threads = 128;     // CTA size
  int ipt = 1;           // instances per thread (1, 2 or 4)
  int slots_shared = 0;  // shared work slots per instance
  int slots_global = 0;  // scratch slots per instance
  int ctas_per_sm = 0;   // filled by plan_occupancy
  int grid = 0;          // persistent grid size
  size_t smem_bytes = 0;
};

// packed words (ccu_isa.h) -> records for `lanes_per_cta` = threads * ipt, padded with X_END to whole chunks
std::vector<XInstr> predecode(const std::vector<uint64_t>& words, int lanes_per_cta);

// dynamic shared memory of a plan: work vector + two program stages + two mbarriers
size_t plan_smem_bytes(const LaunchPlan& plan);

// fills ctas_per_sm/grid/smem_bytes for the current device; returns cudaSuccess or an error
cudaError_t plan_occupancy(LaunchPlan* plan, int device);

// scratch must hold slots_global * grid * threads * ipt doubles; n_records = size of the predecoded program
cudaError_t launch_interp(const LaunchPlan& plan, const XInstr* d_prog, long long n_records, const IoDesc& io,
                          long long N, double* d_scratch, cudaStream_t stream);

}  // namespace ccu
