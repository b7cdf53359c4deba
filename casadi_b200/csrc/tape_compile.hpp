// Host-side "export step": packs an SXFunction tape (ScalarAtomic stream,
// casadi/core/sx_function.hpp:37-44) into the device program of ccu_isa.h.
// Pure C++ (no CUDA) so it can be exercised on a CPU-only box.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ccu {

struct TapeSource {
  long long n_instr = 0;
  const int* op = nullptr;
  const int* i0 = nullptr;
  const int* i1 = nullptr;
  const int* i2 = nullptr;
  const double* d = nullptr;
  long long sz_w = 0;
  std::vector<long long> nnz_in, nnz_out;
};

struct CompileOptions {
  int slots_shared = 0;   // shared-memory work slots per instance available to the allocator (>=4)
  bool use_acc = true;    // forward the previous result in a register (F_ACC / D_NONE)
  int schedule = 0;       // 0 = reference order, 1 = min-cut bisection order (tape_schedule.hpp)
};

struct Program {
  std::vector<uint64_t> words;
  int slots_shared = 0;    // shared slots actually used
  int slots_global = 0;    // scratch slots actually used
  long long n_instr = 0;   // source instructions
  long long flops = 0;     // arithmetic source instructions
  long long max_live = 0;  // peak number of simultaneously live values
  long long spill_loads = 0, spill_stores = 0, remat = 0;
  long long smem_loads = 0, smem_stores = 0;  // shared-memory accesses of the emitted program
};

// The value graph (SSA) recovered from the slot stream: node k is tape instruction k; operands refer to the
// defining node.  Shared by the interpreter's slot allocator and the specialising code generator (jit.cpp).
enum Kind : uint8_t { K_ARITH, K_CONST, K_INPUT, K_OUTPUT };
struct Node {
  uint8_t kind;
  uint8_t dop;          // DevOp (ccu_isa.h)
  int a = -1, b = -1;   // operand value ids (node indices); OUTPUT: a = source value
  int idx = 0, nz = 0;  // INPUT: input index / nonzero; OUTPUT: output index / nonzero
  double c = 0;         // CONST literal
  // inside a re-rolled loop body (tape_reroll.hpp; both are 0 / -1 everywhere else):
  int vary = -1;        // CONST: column of the per-iteration constant table (the literal differs between iterations)
  int step = 0;         // INPUT: the nonzero index advances by `step` per iteration
};
// Validates the tape and builds the graph; `flops` = arithmetic instructions of the tape.  Instructions that
// repeat an earlier one on the same operand values are dropped (value numbering; CCU_CSE=0 keeps them): `removed` = how
// many arithmetic instructions that saved.
bool build_graph(const TapeSource& s, std::vector<Node>* nodes, long long* flops, std::string* err, long long* removed = nullptr);

// Validates the tape (throws nothing; returns false and sets err) and computes the number of
// simultaneously live values -- what an allocation without spills needs.
bool analyse_tape(const TapeSource& src, long long* max_live, long long* flops, std::string* err, long long* removed = nullptr);

bool compile_tape(const TapeSource& src, const CompileOptions& opt, Program* out, std::string* err);

}  // namespace ccu
