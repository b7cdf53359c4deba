/*
 * casadi_cuda.h -- C ABI of libcasadi_cuda.so, the B200 (sm_100a) evaluator behind
 * CasADi's `Function::map(N, "cuda")`.
 *
 * Plain C: pointers and sizes only, no CasADi, torch or CUDA types in any signature
 * (a CUDA stream is passed as void*).  Each entry point names the reference interface
 * it replaces (paths relative to the casadi/casadi 3.7.2 tree).  `casadi_b200/host/cuda_map.cpp`
 * (the `CudaMap` subclass of casadi::Map) and `casadi_b200/capi.py` (ctypes) are the two
 * bindings shipped in this repository; INTEGRATION.md shows the reference-side patch.
 *
 * Conventions
 *   - every function returning int returns 0 on success, non-zero on failure; the message is
 *     available from ccu_last_error() (thread-local).  There is NO CPU fallback: when no CUDA
 *     device is usable the create/eval calls fail.
 *   - ccu_int is casadi_int (long long, casadi/core/casadi_types.hpp:29-38).
 *   - "AoS" is the reference's Map layout: instance i of input j is arg[j][i*nnz_in[j] .. +nnz_in[j])
 *     (casadi/core/map.cpp:149-154, map.hpp:77-82).  "SoA" is [k][i]: element k of instance i at
 *     arg[j][k*N + i] (device-resident fast path, fully coalesced).
 */
#ifndef CASADI_CUDA_H
#define CASADI_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define CCU_EXPORT __declspec(dllexport)
#else
#define CCU_EXPORT __attribute__((visibility("default")))
#endif

typedef long long ccu_int;

#define CCU_ABI_VERSION 3
#define CCU_LAYOUT_AOS 0
#define CCU_LAYOUT_SOA 1
#define CCU_MODE_INTERP 0 /* tape-interpreter kernel (always available)                                  */
#define CCU_MODE_JIT 1    /* tape specialised into straight-line sm_100a kernels by NVRTC at create time */
#define CCU_MODE_AUTO 2   /* JIT when possible, else INTERP (ccu_set_default_mode only)                  */

/* opaque handles */
typedef struct ccu_tape ccu_tape;     /* a compiled SX instruction tape, resident on one device      */
typedef struct ccu_linsol ccu_linsol; /* a symbolic LDL / QR factorisation shared by a whole batch   */
typedef struct ccu_comm ccu_comm;     /* NCCL communicator(s) of this process (reduce_out sums across GPUs) */
typedef struct ccu_multi ccu_multi;   /* one tape replicated on several devices of this process + ccu_comm  */

/* ------------------------------------------------------------------------------------------------
 * library
 * ---------------------------------------------------------------------------------------------- */
CCU_EXPORT int ccu_abi_version(void);
CCU_EXPORT const char* ccu_last_error(void);
/* number of usable CUDA devices (0 when there is none; never an error) */
CCU_EXPORT int ccu_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * SX tape  --  replaces SXFunction::eval (casadi/core/sx_function.cpp:72-127) over the tape
 * `std::vector<ScalarAtomic> algorithm_` (sx_function.hpp:37-44,258), exported through the public
 * accessors Function::n_instructions/instruction_id/_input/_output/_constant (function.hpp:1114-1138).
 *
 * op[k]  : enum Operation value (calculus.hpp:60-218)
 * i0[k]  : destination work slot            (OP_OUTPUT: output index)
 * i1,i2  : source work slots                (OP_INPUT: input index, nonzero index;
 *                                            OP_OUTPUT: i1 = source slot, i2 = nonzero index)
 * d[k]   : value of an OP_CONST instruction
 * sz_w   : work vector length (SXFunction::worksize_)
 * Fails (NULL) for tapes the device cannot evaluate: OP_CALL, OP_PARAMETER (free variables,
 * sx_function.cpp:78-83), OP_PRINTME, unknown opcodes, out-of-range indices.
 * ---------------------------------------------------------------------------------------------- */
CCU_EXPORT ccu_tape* ccu_tape_create(ccu_int n_instr, const int* op, const int* i0, const int* i1,
                                     const int* i2, const double* d, ccu_int sz_w, ccu_int n_in,
                                     const ccu_int* nnz_in, ccu_int n_out, const ccu_int* nnz_out,
                                     int device);
CCU_EXPORT void ccu_tape_destroy(ccu_tape* t);

/* static facts about a compiled tape (all per ONE evaluation) */
typedef struct ccu_tape_info {
  ccu_int n_instr;        /* instructions in the source tape                                         */
  ccu_int n_words;        /* 8-byte words of the packed device program                               */
  ccu_int flops;          /* arithmetic tape instructions (op<OP_CONST minus ASSIGN, +5)  */
  ccu_int bytes_in;       /* 8 * sum nnz_in                                                          */
  ccu_int bytes_out;      /* 8 * sum nnz_out                                                         */
  ccu_int sz_w;           /* work slots of the source tape                                           */
  ccu_int slots_shared;   /* work slots kept in shared memory per instance                           */
  ccu_int slots_global;   /* work slots spilled to the global scratch per instance                   */
  ccu_int threads;        /* CTA size the plan launches                                              */
  ccu_int ipt;            /* instances per thread                                                    */
  ccu_int smem_bytes;     /* dynamic shared memory per CTA                                           */
  ccu_int spill_loads;    /* global-scratch reads per evaluation                                     */
  ccu_int spill_stores;   /* global-scratch writes per evaluation                                    */
  ccu_int max_live;       /* peak number of simultaneously live values of the tape                   */
  ccu_int grid;           /* persistent grid of the interpreter plan                                 */
  ccu_int ctas_per_sm;    /* resident CTAs per SM of the interpreter plan                            */
  ccu_int mode;           /* CCU_MODE_INTERP / CCU_MODE_JIT: what the evaluation calls launch        */
  ccu_int jit_segments;   /* specialised kernels per tile (0 = not built)                            */
  ccu_int jit_scratch_slots; /* cross-segment values per instance                                    */
  ccu_int jit_tile;       /* instances per tile (0 = automatic)                                      */
  ccu_int jit_compile_ms; /* generation + NVRTC + load time                                          */
  ccu_int jit_cross_loads;   /* scratch reads per evaluation                                         */
  ccu_int jit_cross_stores;  /* scratch writes per evaluation                                        */
  ccu_int jit_max_regs;   /* max registers per thread over the segments                              */
  ccu_int jit_cache_hits; /* segments served from the cubin cache                                    */
  ccu_int jit_threads;    /* CTA size of the specialised kernels                                     */
  ccu_int jit_schedule;   /* 0 = reference order, 1 = min-cut bisection order (csrc/tape_schedule.hpp)       */
  ccu_int jit_schedule_ms;/* time spent ordering and cutting the tape                                        */
  ccu_int jit_chained;    /* 1 = the segments are linked into ONE persistent kernel (one launch per evaluation)*/
  ccu_int cse_removed;    /* arithmetic instructions the device does NOT execute: they repeat an earlier instruction on the same
                             operand values (value numbering at tape creation; `flops` stays the reference's count)   */
  ccu_int jit_remat_cloned; /* instructions the specialised kernels RECOMPUTE in a reading segment instead of storing and loading
                             their value (rematerialisation, csrc/tape_schedule.hpp)                                  */
  ccu_int jit_loop_iters; /* re-rolled plan (csrc/tape_reroll.hpp): iterations of the persistent loop kernel (0 = flat plan) */
  ccu_int jit_loop_body;  /* arithmetic instructions of one iteration                                               */
  ccu_int jit_loop_slots; /* slots of the per-CTA loop scratch (0 = the loop state lives in registers)              */
} ccu_tape_info;
CCU_EXPORT int ccu_tape_get_info(const ccu_tape* t, ccu_tape_info* info);

/* Debug/verification access to the packed device program (ISA: casadi_b200/csrc/ccu_isa.h):
 * copies min(cap, n_words) words and returns n_words.  `device` = -1 in ccu_tape_create compiles
 * without a GPU (such a tape can be inspected but never evaluated). */
CCU_EXPORT ccu_int ccu_tape_get_program(const ccu_tape* t, unsigned long long* words, ccu_int cap);

/* Execution mode.  ccu_tape_create honours the environment variable CCU_MODE = interp | jit | auto
 * (default auto: specialise when NVRTC is loadable and the tape compiles, else interpret).  The reference's
 * analogue is the Function option "jit" (casadi/core/function_internal.cpp, options "jit"/"compiler").
 * ccu_tape_set_mode(CCU_MODE_JIT) fails when the specialisation is impossible; there is no CPU path. */
CCU_EXPORT int ccu_tape_set_mode(ccu_tape* t, int mode);
/* Process-wide default for subsequent ccu_tape_create calls: CCU_MODE_INTERP, CCU_MODE_JIT (create fails when
 * the specialisation fails), CCU_MODE_AUTO, or -1 = take it from the environment (the initial state). */
CCU_EXPORT int ccu_set_default_mode(int mode);
/* Tunables of the specialisation: arithmetic instructions per segment, CTA size, __launch_bounds__
 * min-blocks (0 = none), instances per tile (0 = automatic); arguments <= 0 (tile, min_blocks: < 0) keep the
 * current value.  Rebuilds the kernels and selects CCU_MODE_JIT. */
CCU_EXPORT int ccu_tape_set_jit_plan(ccu_tape* t, int seg_instr, int threads, int min_blocks, ccu_int tile);
/* Generated CUDA source of segment `segment` (inspection/tests; works without a GPU): copies at most cap-1
 * characters, returns the full length; segment < 0 returns the number of segments. */
CCU_EXPORT ccu_int ccu_tape_get_jit_source(const ccu_tape* t, ccu_int segment, char* buf, ccu_int cap);
/* Plan of the specialisation for given options, computed on the host (works without a GPU; nothing is compiled or
 * changed): seg_instr <= 0 / schedule < 0 keep the tape's current option.  The SXFunction tape order is the
 * reference's depth-first order (sx_function.cpp:522-540); schedule 1 re-orders it (bit-identical results).
 * stats = {segments, scratch slots, scratch reads per evaluation, scratch writes per evaluation,
 *          largest segment (arithmetic instructions), schedule time in ms,
 *          peak values alive inside one segment (max over segments), the same (mean over segments)} -- of the order and
 * the cuts alone, without the automatic rematerialisation (ccu_tape_jit_remat_stats). */
CCU_EXPORT int ccu_tape_jit_plan_stats(const ccu_tape* t, int seg_instr, int schedule, ccu_int stats[8]);
/* The same plan with rematerialisation (csrc/tape_schedule.hpp: a cross-segment value is recomputed in the reading
 * segment when a minimum cut prices that below `remat` FP64 issue slots per stored + loaded value; 0 = off,
 * < 0 = the tape's current setting).  stats = {recomputed instructions added, instructions dropped from their
 * defining segment, scratch reads per evaluation, scratch writes per evaluation, segments, scratch slots}. */
CCU_EXPORT int ccu_tape_jit_remat_stats(const ccu_tape* t, int seg_instr, int remat, ccu_int stats[6]);
/* Loop re-rolling (csrc/tape_reroll.hpp): the time-stepping loop recovered from the unrolled tape (host only, no GPU).
 * stats = {1 when a loop was found and verified, iterations, arithmetic instructions per iteration, values carried from
 * one iteration to the next, constants that differ between iterations, input nonzeros that advance with the iteration,
 * values of the last iteration read after the loop, arithmetic instructions before the loop}; when no loop is found
 * stats[0] = 0 and ccu_last_error says why. */
CCU_EXPORT int ccu_tape_loop_stats(const ccu_tape* t, ccu_int stats[8]);
/* Sets the rematerialisation price (see above) and rebuilds the specialised kernels. */
CCU_EXPORT int ccu_tape_set_jit_remat(ccu_tape* t, int remat);
/* Compiles the segments of the current plan as relocatable device functions plus the persistent chain kernel and
 * links them for sm_100a (NVRTC + nvJitLink, both dlopen'ed; works without a GPU).  Returns the size of the linked
 * cubin in bytes, -1 on failure.  The reference's analogue: the "jit" option compiling the generated C of a whole
 * Function into one shared object (function_internal.cpp, importer.cpp). */
CCU_EXPORT ccu_int ccu_tape_jit_link_check(const ccu_tape* t);
/* Why the built specialisation runs one kernel per segment instead of the chain ("" when chained / not built). */
CCU_EXPORT const char* ccu_tape_jit_chain_error(const ccu_tape* t);
/* Build check for CPU-only boxes: compiles every kernel of the current plan for sm_100a with NVRTC (no GPU needed) and,
 * when dump_dir is not NULL, writes the cubins there as kernel<k>.cubin.  Returns their total size, -1 on failure. */
CCU_EXPORT ccu_int ccu_tape_jit_compile_check(const ccu_tape* t, const char* dump_dir);
/* Selects the order used by subsequent (re)builds of the specialised kernels and by ccu_tape_get_jit_source. */
CCU_EXPORT int ccu_tape_set_jit_schedule(ccu_tape* t, int schedule);

/* Tunables of the plan (threads per CTA, instances per thread, shared slots); 0 = choose
 * automatically.  Replaces nothing in the reference: Map::create passes an empty Dict (map.cpp:43-47). */
CCU_EXPORT int ccu_tape_set_plan(ccu_tape* t, int threads, int ipt, int slots_shared);

/* ------------------------------------------------------------------------------------------------
 * Map evaluation  --  replaces Map::eval / Map::eval_gen (casadi/core/map.cpp:141-157, 327-334) and
 * OmpMap::eval (map.cpp:340-386).
 *
 * ccu_map_eval_host: arg[j], res[j] are HOST pointers in the reference's AoS layout; arg[j]==NULL
 * reads as zeros (sx_function.cpp:116), res[j]==NULL is not computed (:117).  Does H2D, the kernel
 * and D2H; returns after the results are in res.
 * ---------------------------------------------------------------------------------------------- */
CCU_EXPORT int ccu_map_eval_host(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res);

/* Same with DEVICE pointers (on the tape's device) and an explicit layout; asynchronous on `stream`
 * (a cudaStream_t, NULL = the legacy default stream).  This is the call the roofline is measured on. */
CCU_EXPORT int ccu_map_eval_device(ccu_tape* t, ccu_int N, const double* const* d_arg,
                                   double* const* d_res, int layout, void* stream);

/* Map with reductions -- replaces HorzRepmat/HorzRepsum around a Map (function.cpp:797-818,
 * repmat.cpp:44-50,127-135) and MapSum::eval_gen (mapsum.cpp:154-186).
 * reduce_in[j]!=0 : input j is ONE instance (nnz_in[j] doubles) broadcast to all N evaluations.
 * reduce_out[j]!=0: output j is the sum over the N evaluations (nnz_out[j] doubles).
 * The device sum is a fixed-shape pairwise tree (independent of launch geometry and GPU count);
 * it differs from the reference's sequential sum by rounding only. Pointers as in ccu_map_eval_host. */
CCU_EXPORT int ccu_map_eval_reduce_host(ccu_tape* t, ccu_int N, const double* const* arg,
                                        double* const* res, const int* reduce_in, const int* reduce_out);
CCU_EXPORT int ccu_map_eval_reduce_device(ccu_tape* t, ccu_int N, const double* const* d_arg,
                                          double* const* d_res, const int* reduce_in,
                                          const int* reduce_out, int layout, void* stream);

/* Multi-GPU: the instances [i0, i0+n) of a batch of N_global are evaluated by this tape's device (instances are
 * independent, map.cpp:147-155: contiguous shards, no exchange).  d_arg / d_res point at the SHARD's data (its
 * first instance at index 0; SoA leading dimension n).  For reduce_out outputs the level-0 sums of the shard's
 * 1024-instance blocks are written at their GLOBAL block positions into d_part[j], a device vector of
 * ceil(N_global/1024)*nnz_out[j] doubles that the caller zero-initialises; i0 must be a multiple of 1024.
 * Summing d_part over the ranks (disjoint supports: the sum is exact -- one NCCL all-reduce) and then calling
 * ccu_reduce_tree_device gives a result that is bit-identical for every number of GPUs, and to
 * ccu_map_eval_reduce_device on one GPU.  d_part is overwritten by ccu_reduce_tree_device. */
CCU_EXPORT int ccu_map_eval_shard_device(ccu_tape* t, ccu_int N_global, ccu_int i0, ccu_int n,
                                         const double* const* d_arg, double* const* d_res, const int* reduce_in,
                                         const int* reduce_out, double* const* d_part, int layout, void* stream);
CCU_EXPORT int ccu_reduce_tree_device(int device, double* d_part, ccu_int N_global, ccu_int nnz, double* d_out,
                                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tape builder  --  records scalar operations, and whole runtime algorithms traced over a sparsity pattern
 * shared by the batch, into the tape format above.  Replaces, for mapped functions, the per-instance calls
 *   casadi_ldl / casadi_ldl_solve   casadi/core/runtime/casadi_ldl.hpp:25-109  (LinsolLdl::nfact/solve,
 *                                   casadi/solvers/linsol_ldl.cpp:119-132)
 *   casadi_qr / casadi_qr_solve /   casadi/core/runtime/casadi_qr.hpp:24-227   (LinsolQr::nfact/solve,
 *   casadi_qr_singular              casadi/solvers/linsol_qr.cpp:126-180)
 *   casadi_mtimes                   casadi/core/runtime/casadi_mtimes.hpp:22-75 (Multiplication::eval)
 * with the same floating-point operations in the same order.  Patterns are the reference's compressed CCS
 * vectors [nrow, ncol, colind[ncol+1], row[nnz]] (Sparsity::operator const casadi_int*, sparsity.cpp:1734);
 * values are handles returned by the builder.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ccu_builder ccu_builder;
CCU_EXPORT ccu_builder* ccu_builder_create(void);
CCU_EXPORT void ccu_builder_destroy(ccu_builder* b);
CCU_EXPORT ccu_int ccu_builder_const(ccu_builder* b, double c);                 /* -> handle            */
CCU_EXPORT ccu_int ccu_builder_input(ccu_builder* b, ccu_int idx, ccu_int nz);  /* -> handle of arg[idx][nz] */
/* op = enum Operation (calculus.hpp:60-218); y is ignored for unary operations.  -> handle, -1 on error */
CCU_EXPORT ccu_int ccu_builder_op(ccu_builder* b, int op, ccu_int x, ccu_int y);
CCU_EXPORT int ccu_builder_output(ccu_builder* b, ccu_int idx, ccu_int nz, ccu_int v); /* res[idx][nz] = v */
/* a: nnz(A) handles; x: n*nrhs handles, right-hand sides on entry, solutions on return.
 * zero_pivots (optional): handle of the number of zeros in D (LinsolLdl::nfact warns, linsol_ldl.cpp:122-124). */
CCU_EXPORT int ccu_builder_ldl(ccu_builder* b, const ccu_int* sp_a, const ccu_int* sp_lt, const ccu_int* p,
                               const ccu_int* a, ccu_int* x, ccu_int nrhs, ccu_int* zero_pivots);
/* nullity (optional): handle of the number of |R_cc| < eps (LinsolQr::nfact fails when > 0, linsol_qr.cpp:146-163) */
CCU_EXPORT int ccu_builder_qr(ccu_builder* b, const ccu_int* sp_a, const ccu_int* sp_v, const ccu_int* sp_r,
                              const ccu_int* prinv, const ccu_int* pc, const ccu_int* a, ccu_int* x, ccu_int nrhs,
                              int tr, double eps, ccu_int* nullity);
/* z += x*y (handles in z are replaced) */
CCU_EXPORT int ccu_builder_mtimes(ccu_builder* b, const ccu_int* x, const ccu_int* sp_x, const ccu_int* y,
                                  const ccu_int* sp_y, ccu_int* z, const ccu_int* sp_z);
/* c != 0 ? a : b with the selected operand's bits (sign of zero, NaN) preserved; recorded as seven operations of the
 * reference's set (if_else_zero, not, copysign, add, mul) -> handle */
CCU_EXPORT ccu_int ccu_builder_select(ccu_builder* b, ccu_int c, ccu_int x, ccu_int y);
/* The recorded program in the reference's tape layout (the arrays ccu_tape_create takes): copies up to `cap`
 * instructions into op / i0 / i1 / i2 / d (each may be NULL), stores the work-vector size in *sz_w (may be NULL) and
 * returns the instruction count.  Needs no device: what a host-side check evaluates instance by instance. */
CCU_EXPORT ccu_int ccu_builder_export(const ccu_builder* b, int* op, int* i0, int* i1, int* i2, double* d, ccu_int cap,
                                      ccu_int* sz_w);
/* compile what has been recorded; the builder can be destroyed afterwards */
CCU_EXPORT ccu_tape* ccu_builder_finish(ccu_builder* b, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out,
                                        const ccu_int* nnz_out, int device);

/* ------------------------------------------------------------------------------------------------
 * Batched linear solves with shared sparsity: factorise and solve N systems A_i x_i = b_i.
 * A: N x nnz(A) values (AoS, CCS order), B/X: N x (n*nrhs).  n_flagged / d_flagged receive the number of
 * instances with a zero pivot in D (LDL) or with a numerically singular R (QR, |R_cc| < eps).
 * ---------------------------------------------------------------------------------------------- */
CCU_EXPORT ccu_linsol* ccu_ldl_create(const ccu_int* sp_a, const ccu_int* sp_lt, const ccu_int* p, ccu_int nrhs,
                                      int device);
CCU_EXPORT ccu_linsol* ccu_qr_create(const ccu_int* sp_a, const ccu_int* sp_v, const ccu_int* sp_r,
                                     const ccu_int* prinv, const ccu_int* pc, ccu_int nrhs, int tr, double eps,
                                     int device);
CCU_EXPORT void ccu_linsol_destroy(ccu_linsol* ls);
CCU_EXPORT ccu_tape* ccu_linsol_tape(ccu_linsol* ls); /* the traced tape (info, mode, plan); owned by ls */
CCU_EXPORT int ccu_linsol_solve_host(ccu_linsol* ls, ccu_int N, const double* A, const double* B, double* X,
                                     ccu_int* n_flagged);
CCU_EXPORT int ccu_linsol_solve_device(ccu_linsol* ls, ccu_int N, const double* d_A, const double* d_B, double* d_X,
                                       double* d_flagged, int layout, void* stream);

/* Time (ms) of the most recent kernel launch sequence of this tape, measured with CUDA events on
 * the launching stream (synchronises).  FStats analogue (casadi/core/timing.hpp:47-98). */
CCU_EXPORT int ccu_tape_last_kernel_ms(ccu_tape* t, double* ms);
/* Phase times of the most recent ccu_map_eval_host / ccu_map_eval_reduce_host call on this tape -- the split the
 * reference's FStats would record around the three phases (casadi/core/timing.hpp:47-98,
 * function_internal.cpp:986-1011): stats[0] = H2D, [1] = kernels (layout + tape + reduction), [2] = D2H, each the
 * device time summed over the chunks (the phases overlap, so they do not add up to the wall time), [3] = time the
 * calling thread spent copying between pageable caller memory and the pinned staging, [4] = wall time of the call,
 * all in ms; [5] = bytes that went through the pinned staging (0 when every caller buffer was page-locked). */
CCU_EXPORT int ccu_tape_last_eval_stats(const ccu_tape* t, double stats[6]);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
CCU_EXPORT ccu_int ccu_launch_count(void);
/* Measured FP64 non-FMA issue rate (DADD/s) of `device`: the FP64 denominator of the roofline (SURVEY 8d;
 * contraction is off by contract, so one tape instruction is at best one DADD).  Runs a ~3 ms microbenchmark. */
CCU_EXPORT int ccu_fp64_issue_rate(int device, double* ops_per_s);
/* Self test of the branch-free division / sin / cos fast paths the specialised kernels use (csrc/ccu_ops.cuh) against
 * the plain operators (div.rn.f64, sin, cos, sincos) on `n` generated operand pairs: raw bit patterns, moderate
 * magnitudes, trig arguments across the 2^31 fast-path limit and special values.  counts[0] = results that differ in
 * a bit although the fast path did not flag the operands (must be 0), counts[1] = flagged (re-evaluated by the plain
 * operator in a kernel), counts[2] = checks.  Restates nothing of the reference: parity infrastructure of this library. */
CCU_EXPORT int ccu_selftest_fastops(int device, long long n, unsigned long long seed, unsigned long long counts[3]);
/* Self test of the host path's staging copy (the multi-threaded copy with non-temporal stores between a caller's
 * pageable buffer and the pinned staging, csrc/hostcopy.cpp): copies `bytes` bytes between two heap buffers offset by
 * dst_misalign / src_misalign bytes and compares with the source and the guard bytes around the destination.  Needs
 * no GPU.  Returns 0 when identical; *gb_per_s (may be NULL) receives the copy rate. */
CCU_EXPORT int ccu_selftest_host_copy(long long bytes, int dst_misalign, int src_misalign, double* gb_per_s);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU.  Instances are independent (map.cpp:147-155): the batch is cut into contiguous shards of
 * whole 1024-instance reduction blocks, one per GPU, the tape is replicated, and NOTHING is exchanged for a plain
 * map.  Only reduce_out communicates (HorzRepsum / MapSum sums, repmat.cpp:127-135, mapsum.cpp:170-184): every GPU
 * writes the level-0 sums of its blocks at their global rows of a zero vector, ONE NCCL all-reduce over NVLink /
 * NVSwitch merges the vectors (on the 64-bit patterns: the supports are disjoint, so the merge is exact down to -0.0
 * and NaN payloads) and the level-1 tree runs on the merged vector -- the sums have the bits of a single-GPU run for
 * any number of GPUs.  libnccl is bound at run time (CCU_NCCL_LIB, libnccl.so.2); single-GPU use never loads it.
 * ---------------------------------------------------------------------------------------------- */
/* 1 when libnccl can be loaded in this process (else 0 and ccu_last_error says why) */
CCU_EXPORT int ccu_comm_available(void);
CCU_EXPORT int ccu_comm_nccl_version(void);
/* one process driving several devices: ncclCommInitAll over devices[0..n) (devices == NULL: 0..n-1) */
CCU_EXPORT ccu_comm* ccu_comm_create_all(int n_devices, const int* devices);
/* one rank of a multi-process job: id from ccu_comm_unique_id on one rank, distributed by the host (ncclCommInitRank) */
CCU_EXPORT int ccu_comm_unique_id(unsigned char id[128]);
CCU_EXPORT ccu_comm* ccu_comm_create_rank(const unsigned char id[128], int rank, int n_ranks, int device);
CCU_EXPORT void ccu_comm_destroy(ccu_comm* c);
CCU_EXPORT int ccu_comm_size(const ccu_comm* c);
/* In-place all-reduce of the zero-padded block-sum vectors d_part[k][0..count) -- one per LOCAL device of the
 * communicator, on streams[k] (NULL: default streams) -- as produced by ccu_map_eval_shard_device; follow with
 * ccu_reduce_tree_device.  Asynchronous on the streams. */
CCU_EXPORT int ccu_comm_allreduce_block_sums(ccu_comm* c, double* const* d_part, ccu_int count, void* const* streams);

/* `f.map(N, "cuda")` over several devices of ONE process (what CudaMap builds for CASADI_CUDA_DEVICES=all|0,1,..):
 * replicas of the tape of ccu_tape_create / ccu_builder_finish on devices[0..n) and, for n > 1, their communicator. */
CCU_EXPORT ccu_multi* ccu_multi_create(ccu_int n_instr, const int* op, const int* i0, const int* i1, const int* i2,
                                       const double* d, ccu_int sz_w, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out,
                                       const ccu_int* nnz_out, int n_devices, const int* devices);
CCU_EXPORT ccu_multi* ccu_builder_finish_multi(ccu_builder* b, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out,
                                               const ccu_int* nnz_out, int n_devices, const int* devices);
CCU_EXPORT void ccu_multi_destroy(ccu_multi* m);
CCU_EXPORT int ccu_multi_size(const ccu_multi* m);
CCU_EXPORT ccu_tape* ccu_multi_tape(ccu_multi* m, int k); /* replica k (owned by m) */
/* ccu_map_eval_reduce_host over all devices of m: device g evaluates instances [g*N/G, (g+1)*N/G) (whole reduction
 * blocks) through its own chunked host pipeline, concurrently; reduced outputs are combined by NCCL as described
 * above.  reduce_in / reduce_out may be NULL (plain map).  Same arguments and error behaviour as the single-device call. */
CCU_EXPORT int ccu_multi_eval_host(ccu_multi* m, ccu_int N, const double* const* arg, double* const* res,
                                   const int* reduce_in, const int* reduce_out);
/* The same with PIECE-major caller buffers: in_groups[j] = G > 1 says that instance k of input j is G pieces of
 * nnz_in(j)/G doubles and piece d lies at arg[j] + (d*N + k)*nnz_in(j)/G (out_groups likewise).  This is the layout
 * of the nfwd / nadj seed and sensitivity blocks of a derivative map; the reference converts it to the instance-major
 * layout of df.map(n) with GetNonzeros column permutations around the call (Map::get_forward / get_reverse,
 * map.cpp:231-264, 285-318) -- here the permutation is folded into the chunk copies.  NULL = ungrouped. */
CCU_EXPORT int ccu_multi_eval_host_grouped(ccu_multi* m, ccu_int N, const double* const* arg, double* const* res,
                                           const int* reduce_in, const int* reduce_out, const int* in_groups,
                                           const int* out_groups);

/* ------------------------------------------------------------------------------------------------
 * Device memory helpers (so a C or C++ host needs no CUDA headers)
 * ---------------------------------------------------------------------------------------------- */
CCU_EXPORT int ccu_set_device(int device);
CCU_EXPORT void* ccu_malloc(ccu_int bytes);
CCU_EXPORT int ccu_free(void* p);
CCU_EXPORT void* ccu_malloc_host(ccu_int bytes); /* pinned */
CCU_EXPORT int ccu_free_host(void* p);
/* Page-lock a caller's buffer IN PLACE (cudaHostRegister, portable across devices): ccu_map_eval_host and friends then
 * let the DMA engines read / write it directly instead of copying every byte through the library's pinned staging
 * (quadrotor step end to end: +45 % when the buffers are long-lived).  The buffer must stay allocated until
 * ccu_host_unregister; a buffer that is page-locked already (cudaHostAlloc, the caller's own cudaHostRegister) needs
 * nothing.  CCU_HOST_REGISTER=1 in the environment does the same automatically for every mapped buffer of >= 1 MiB the
 * first time an evaluation sees it, and un-registers when the tape is destroyed -- meant for the reference's buffer API
 * (Function::operator()(arg,res,iw,w,mem), function.hpp) called repeatedly with the same buffers, NOT for callers that
 * free and re-allocate buffers between calls.  ccu_host_registered_count: registrations currently held. */
CCU_EXPORT int ccu_host_register(const void* p, ccu_int bytes);
CCU_EXPORT int ccu_host_unregister(const void* p);
CCU_EXPORT int ccu_host_registered_count(void);
CCU_EXPORT int ccu_memcpy_h2d(void* dst, const void* src, ccu_int bytes, void* stream);
CCU_EXPORT int ccu_memcpy_d2h(void* dst, const void* src, ccu_int bytes, void* stream);
CCU_EXPORT int ccu_stream_sync(void* stream);
CCU_EXPORT int ccu_device_sync(void);

#ifdef __cplusplus
}
#endif
#endif /* CASADI_CUDA_H */
