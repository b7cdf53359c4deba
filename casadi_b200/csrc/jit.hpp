// Tape specialisation ("jit" mode): an SX tape is turned into straight-line sm_100a kernels whose work
// vector lives in REGISTERS (static register indices are the one thing an interpreter cannot have --
// SURVEY 7 "central risk" ii), compiled at tape-creation time by NVRTC with --fmad=false.
//
// This is the analogue of the reference's own `jit` option (FunctionInternal option "jit",
// casadi/core/function_internal.cpp; code generator SXFunction::codegen_body, sx_function.cpp:344-434),
// not a replacement of the interpreter: when NVRTC is not loadable, or a tape fails to compile, the tape
// runs on the interpreter kernel (interp.cu).  Never a CPU path.
//
// Long tapes are cut into SEGMENTS of ~seg_instr arithmetic instructions, one kernel each, so that ptxas
// sees bounded basic blocks (compile time is superlinear in block length) and segments compile in parallel.
// Values that cross a segment boundary travel through a global scratch laid out [slot][instance of the
// tile] (coalesced); constants and inputs are re-materialised.  The batch is processed in tiles so the
// scratch is sized by the tile, not by N.  The ORDER of operations and every operand pairing is the tape's:
// each value is computed by the same IEEE operation from the same operand values as in SXFunction::eval
// (sx_function.cpp:111-124).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "interp.cuh"
#include "tape_compile.hpp"
#include "tape_schedule.hpp"

namespace ccu {

struct JitOptions {
  int seg_instr = 0;      // arithmetic instructions per segment; 0 = automatic (jit_resolve)
  long long seg_weight = -1;  // estimated SASS instructions per segment (instruction-cache bound): 0 = none, -1 = automatic
  int sincos = 1;         // sin(x) and cos(x) of one operand inside a segment become one sincos()
  int fastops = 1;        // divisions and sin/cos run on the branch-free fast paths of ccu_ops.cuh; a thread whose operands
                          // leave their range re-evaluates the segment with the plain operators (bit-identical either way)
  std::vector<std::pair<double, double>> div_recip;  // (constant divisor, its refined reciprocal as computed on the
                                                     // device): filled by jit_build, empty = divide in the general form
  int remat = -1;         // rematerialisation: FP64 issue slots one cross-segment value is worth (tape_schedule.hpp RematOptions);
                          // 0 = off, -1 = automatic
  int roll = -1;          // re-roll the time-stepping loop of the tape (tape_reroll.hpp) and run it as one persistent loop kernel:
                          // 0 = off, 1 = whenever a loop is found, -1 = automatic
  int roll_registers = 1; // a small loop state lives in registers (0: always in the per-CTA loop scratch)
  int interleave = 0;     // > 0: inside a segment, re-order windows of this many instructions level by level (ILP)
  int schedule = 1;       // 0 = reference order, fixed-length segments; 1 = min-cut bisection (tape_schedule.hpp)
  int threads = 0;        // CTA size; 0 = automatic
  int min_blocks = -1;    // __launch_bounds__ second argument (resident CTAs per SM, bounds the registers); 0 = none, -1 = automatic
  int load_batch = 32;    // cross-segment live-ins read straight from global memory are loaded in groups of this many
  int scratch_block = 128; // consecutive instances that share a scratch block [block][slot][instance] (capped by the CTA)
  int ring_inputs = 1;    // inputs go through the ring as well (0 = batched / direct loads)
  int chain = 0;          // persistent chain kernel (all segments linked into one kernel by nvJitLink): 0 = off (default: it
                          // keeps the scratch in L2 but streams 5.7 MB of code per 128 instances through the instruction
                          // cache: 2.1e7 vs 5.8e7 evals/s on the quadrotor Jacobian), -1 = when possible, 1 = required
  int ring = -1;          // rows of the cp.async ring that prefetches a segment's live-ins into shared memory (0 = off, -1 = automatic: 40)
  int prefetch = 0;       // 1 / 2: prefetch.global.L2 / .L1 of every live-in (and dense inputs) at kernel entry
  int stage = 0;          // live-ins per segment staged in shared memory by TMA bulk copies: -1 = as many as fit, 0 = off
  int spill = -2;         // private shared-memory rows for values that do not fit the registers: -1 = what is left, 0 = off,
                          // -2 = automatic: what is left when the launch bounds allow >= 200 registers, else off
  int reg_values = 0;     // doubles planned in registers (0 = from the launch bounds)
  int compile_threads = 0;  // 0 = hardware concurrency (max 32)
  long long tile = 0;     // instances per tile (0 = automatic)
  int streams = 1;        // tiles in flight at once (each on its own stream and scratch region)
  int iobase = 1;         // flat kernels compute the per-instance base address and the byte stride of every operand once at entry
                          // (an access is base + k * stride) instead of i * si + k * sk in 64 bits per access
  int zigzag = 1;         // odd kernels of the chain walk the tile's CTAs in reverse order (L2 reuse across kernels)
  std::string cache_dir;  // compiled cubins are cached here ("" = $CCU_JIT_CACHE or ~/.cache/casadi_cuda)
};

struct JitProgram {
  std::vector<cudaLibrary_t> libs;
  std::vector<cudaKernel_t> kernels;   // one per segment (empty when chained)
  cudaKernel_t chain = nullptr;        // the persistent chain kernel, when the segments could be linked
  int chain_smem = 0;
  int chain_grid = 0;                  // resident CTAs (filled by the caller from the occupancy of `chain`)
  int segments = 0;
  std::string chain_error;             // why the plan is not chained (when it is not)
  std::vector<int> smem_bytes;  // dynamic shared memory of each kernel (staged live-ins + mbarriers)
  // re-rolled plan (tape_reroll.hpp): kernels = [before-loop segments..., ONE persistent loop kernel, after-loop segments...]
  int loop_kernel = -1;         // index of the loop kernel in `kernels` (-1 = flat plan)
  int loop_iters = 0, loop_body = 0;  // iterations, arithmetic instructions per iteration
  int loop_slots = 0;           // slots of the per-CTA loop scratch (0 = the loop state lives in registers)
  int loop_grid = 0;            // resident CTAs of the loop kernel (filled at build from its occupancy)
  double* d_loop = nullptr;     // loop scratch: loop_slots * threads * loop_grid doubles
  int threads = 128;
  int scratch_slots = 0;        // cross-segment values alive at once (per instance)
  long long tile = 0;           // instances per tile (0 = whole batch in one tile)
  int streams = 1;              // tiles in flight at once
  bool zigzag = true;
  std::vector<cudaStream_t> side;    // created on first use
  std::vector<cudaEvent_t> side_done;
  cudaEvent_t fork = nullptr;
  long long cross_loads = 0;    // scratch reads per evaluation
  long long cross_stores = 0;   // scratch writes per evaluation
  long long smem_moves = 0;     // shared-memory spill stores + reloads per evaluation
  long long remat_cloned = 0;   // instructions recomputed in a reading segment instead of being stored + loaded
  int max_regs = 0;             // max registers per thread over the segments
  int cache_hits = 0;
  double compile_ms = 0;
  double schedule_ms = 0;       // time spent ordering / cutting the tape (tape_schedule.hpp)
};

// Fills in the automatic fields from the tape's size (B200 sweeps, profiles/r1_sweep_sched.jsonl): a tape of up to
// 8000 arithmetic instructions is ONE kernel with 2 x 256 threads per SM (<= 128 registers; the minimum cuts make
// its work vector fit); longer tapes are cut every ~2000 instructions and run 2 x 128 threads per SM with up to 255
// registers, because their segments hold 100-200 values alive at once.
JitOptions jit_resolve(const JitOptions& opt, long long flops, const TapeSource* src = nullptr);

// true when libnvrtc can be loaded in this process
bool jit_available(std::string* why);

// generate + compile + load.  `device` must be the current device.
bool jit_build(const TapeSource& src, const JitOptions& opt, int device, JitProgram* out, std::string* err);

// generated CUDA source of every segment (inspection / tests; no GPU or NVRTC needed)
bool jit_generate(const TapeSource& src, const JitOptions& opt, std::vector<std::string>* sources,
                  JitProgram* plan, std::string* err);

// the plan alone (host only): segments, scratch slots and traffic for the given options
struct JitPlanStats {
  long long segments = 0, scratch_slots = 0, cross_loads = 0, cross_stores = 0, max_segment = 0;
  long long max_live = 0;  // peak number of values alive inside one segment, maximum over the segments
  double mean_live = 0;    // the same, mean over the segments
  double schedule_ms = 0;
  long long remat_cloned = 0, remat_dropped = 0;
};
bool jit_plan_stats(const TapeSource& src, const JitOptions& opt, JitPlanStats* out, std::string* err);

// All segments as relocatable device functions + the persistent chain kernel, compiled by NVRTC and linked by
// nvJitLink into one cubin for `arch` ("sm_100a").  Host only: works without a GPU.
bool jit_link_chain(const std::vector<std::string>& sources, const JitOptions& opt, const std::string& arch, std::string* image,
                    int* cache_hits, std::string* err);

// Compiles every kernel of the plan for `arch` with NVRTC (no GPU needed) -- a build check for CPU-only boxes; the cubins
// are written to `dump_dir` (kernel<k>.cubin) when it is non-empty.  Returns the total size of the cubins, -1 on failure.
long long jit_compile_check(const TapeSource& src, const JitOptions& opt, const std::string& arch, const std::string& dump_dir,
                            std::string* err);

// tile size actually used for a batch of N
long long jit_tile_for(const JitProgram& p, long long N, int sms);

// scratch must hold scratch_slots * jit_tile_for(N) doubles
cudaError_t jit_launch(JitProgram& p, const IoDesc& io, long long N, double* scratch, long long tile,
                       cudaStream_t stream, long long* launches);

void jit_destroy(JitProgram* p);

}  // namespace ccu
