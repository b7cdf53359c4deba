// Tape scheduling: a locality-preserving topological order of the value graph and its cut into segments.
//
// The reference emits a tape depth-first from output 0 (SXFunction::init, casadi/core/sx_function.cpp:522-540,
// XFunction::sort_depth_first x_function.hpp:411-438).  For AD products of time-stepping models that order is
// "row-major over steps": values of every step stay alive while the outputs are produced one after the other, so the
// number of simultaneously live values -- what a device thread has to keep off-chip -- grows with the horizon
// (quadrotor RK4x20: 590 for F, 2 625 for its Jacobian).  Any topological order computes bit-identical results
// (every value is still produced by the same IEEE operation from the same operand values), so the device is free to
// pick one with small cuts:
//
//   recursive bisection -- a piece of the graph is split into a predecessor-closed part D and the rest so that the
//   number of values live across the split is MINIMAL (a minimum s-t cut: value v costs 1 when v is in D and one of
//   its consumers is not; closure constraints are infinite edges).  Balance comes from pinning the first/last
//   `pin_frac` of a topological order of the piece to D / not-D; two orders are tried (the inherited one, which is
//   component-major, and ASAP levels, which are time-major) and the cheaper cut is kept.  Recursion continues down
//   to `min_piece` nodes, which also orders the inside of a segment for register locality.
//
// Segments (one specialised kernel each, jit.cpp) are the maximal pieces of the recursion with at most `seg_instr`
// arithmetic instructions; adjacent small ones are merged.  Pure C++, exercised on CPU-only boxes.
#pragma once
#include <string>
#include <vector>

#include "tape_compile.hpp"

namespace ccu {

struct ScheduleOptions {
  int method = 1;         // 0 = reference order cut every seg_instr arithmetic instructions; 1 = min-cut bisection
  int seg_instr = 800;    // arithmetic instructions per segment (upper bound for method 1)
  long long seg_weight = 0;  // upper bound on a segment's estimated SASS instructions (method 1; 0 = none): kernels
                             // have to fit the instruction cache
  int min_piece = 48;     // pieces of at most this many nodes are not split further
  int max_nodes = 600000; // longer tapes are not re-ordered (method 0)
  double pin_frac = 0.1;  // fraction of a piece pinned to either side of a cut
  int tie_sincos = 1;     // sin(x) and cos(x) of one operand are kept in the same piece (one fused sincos in the kernel)
};

struct Schedule {
  std::vector<int> order;      // new position -> node index; a valid topological order of ALL nodes
  std::vector<int> seg_begin;  // new positions where segments start, plus the sentinel n
  long long cuts = 0;          // minimum cuts computed (diagnostic)
  double ms = 0;               // time spent
};

bool schedule_tape(const std::vector<Node>& nodes, const ScheduleOptions& opt, Schedule* out, std::string* err);

// nodes renumbered to the schedule's order (operand ids rewritten)
void permute_nodes(const std::vector<Node>& in, const std::vector<int>& order, std::vector<Node>* out);

// values read by a later segment than the one defining them: (distinct (value, reading segment) pairs, distinct values)
void cross_traffic(const std::vector<Node>& nodes, const std::vector<int>& seg_begin, long long* loads, long long* stores);

// Rematerialisation (recompute instead of store + load).  A value that crosses a segment boundary costs a store in the
// segment that defines it and a load in every segment that reads it -- 16 bytes of HBM traffic per instance.  For each
// segment a minimum cut decides which of its live-ins (and their ancestors) are RECOMPUTED inside the segment from
// values that are loaded there anyway: a recomputed instruction costs its FP64 issue slots, a loaded value `load_cost`
// of them, an input half that.  Reverse-mode tapes are the case it is made for: the backward sweep of an RK step reads
// ~250 primal intermediates that are all functions of the 12 states at the start of the step (checkpoint the state,
// recompute the step).  Recomputed instructions are the same IEEE operations on the same operand values: bit-identical.
// Values whose every reader now recomputes them are dropped from the segment that used to define (and store) them.
struct RematOptions {
  int load_cost = 12;             // FP64 issue slots one cross-segment value is worth (0 = no rematerialisation)
  int max_candidates = 40000;     // ancestors considered per segment (nearest first)
  long long max_segment_weight = 0; // estimated SASS instructions a segment may hold with its recomputed instructions (0 = no bound)
};
struct RematStats { long long cloned = 0, dropped = 0; };
// returns the number of recomputed instructions added (nodes / seg_begin are rewritten when it is > 0)
long long rematerialise(std::vector<Node>* nodes, std::vector<int>* seg_begin, const RematOptions& opt, RematStats* stats = nullptr);

}  // namespace ccu
