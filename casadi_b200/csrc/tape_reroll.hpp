// Loop re-rolling: recover the time-stepping loop of an unrolled tape.
//
// The reference keeps a T-step integrator (Function::mapaccum / fold, function.cpp:692-746) as a tower of calls when
// it is an MX function, but as soon as it is expanded -- or written as an SX function in the first place, or
// differentiated with SX -- the tape is T nearly identical copies of one step: 20 x 284 instructions for the quadrotor
// integrator, 18 x 3 463 for its Jacobian.  Executed flat, every kernel boundary moves the carried state through HBM
// and every step is separate code.  This pass finds the repetition in the value graph alone, so that the specialising
// code generator (jit.cpp) can emit ONE loop body and iterate it on the device with the carried state on chip.
//
// Method: after value numbering, every output is paired with its counterpart one step earlier (a deepest ancestor with
// the same shallow structural hash, verified by propagating the pairing through the operands); the pairing `pi` is closed
// downwards (operands) and upwards (through the value-numbering table); the orbit of the outputs under pi gives the
// states S_0 (outputs), S_1, S_2, ...; a node belongs to the first state that needs it.  The longest run of consecutive
// "bodies" that pi maps onto each other one-to-one is the loop.  build_loop_template then VERIFIES, operand by operand,
// that iterating the body reproduces the tape exactly (same operation on the same operand values at every position of
// every iteration) -- a tape that does not pass is simply executed flat.  Pure C++, exercised on CPU-only boxes.
#pragma once
#include <string>
#include <vector>

#include "tape_compile.hpp"

namespace ccu {

struct Roll {
  bool found = false;
  std::string why;                   // diagnostic when no loop was found
  int iters = 0;                     // K >= 3
  int body = 0;                      // arithmetic instructions per iteration
  std::vector<std::vector<int>> at;  // at[t][p] = node at body position p of iteration t (positions: a topological order)
  std::vector<int> where;            // node -> -1 = before the loop (leaves too), t in [0, K) = iteration, K = after the loop
  std::vector<int> pos;              // node -> body position, -1 outside the loop
};

// `min_iters`: shortest loop worth reporting
bool find_loop(const std::vector<Node>& nodes, Roll* out, int min_iters = 3);

struct LoopOperand {
  // 0 = same iteration (ref = position)          1 = previous iteration (ref = position; iteration 0 reads `entry`)
  // 2 = the same node for every iteration (ref = node: a value computed before the loop, a constant or an input)
  // 3 = a constant that differs between iterations (ref = column of ctab)
  // 4 = an input nonzero that advances with the iteration (ref = index into affine)
  int kind = -1, ref = -1;
};
struct LoopTemplate {
  int K = 0, B = 0;
  std::vector<uint8_t> dop;
  std::vector<LoopOperand> a, b;     // b.kind = -1 for unary operations
  std::vector<int> carried;          // positions whose value the next iteration reads
  std::vector<int> entry;            // per carried position: the node (outside the loop) iteration 0 reads in its place
  std::vector<std::vector<double>> ctab;  // [column][iteration]
  struct Affine { int idx, nz0, stride; };
  std::vector<Affine> affine;
  std::vector<int> exit_pos;         // positions of the LAST iteration that are read after the loop
};
// Verifies the loop and extracts its template; false (with `why`) when iterating the body would not reproduce the tape.
bool build_loop_template(const std::vector<Node>& nodes, const Roll& roll, LoopTemplate* out, std::string* why);

}  // namespace ccu
