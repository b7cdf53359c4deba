// Tape compiler: SXFunction tape -> device program (see ccu_isa.h, tape_compile.hpp).
//
// The reference tape (SXFunction::init, casadi/core/sx_function.cpp:476-841) addresses a work
// vector of sz_w doubles with slot reuse.  The device keeps the work vector in shared memory, so
// the compiler (1) recovers the value graph (SSA) from the slot stream, (2) re-allocates values to
// at most `slots_shared` shared slots with furthest-next-use eviction, inserting SPILL/FILL moves to
// a per-instance global scratch and re-materialising constants/inputs, and (3) forwards a value
// consumed only by the next instruction through a register (F_ACC/D_NONE).  Every operand pairing is
// preserved: each value is computed by the same operation from the same operand values as in
// SXFunction::eval (sx_function.cpp:111-124), so results are bit-identical whatever the allocation and
// whatever topological order the instructions are issued in (CompileOptions::schedule).
#include "tape_compile.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>
#include <unordered_map>

#include "ccu_isa.h"
#include "tape_schedule.hpp"

namespace ccu {
namespace {

// enum Operation (casadi/core/calculus.hpp:60-218)
enum RefOp {
  R_ASSIGN = 0, R_ADD, R_SUB, R_MUL, R_DIV, R_NEG, R_EXP, R_LOG, R_POW, R_CONSTPOW, R_SQRT, R_SQ, R_TWICE,
  R_SIN, R_COS, R_TAN, R_ASIN, R_ACOS, R_ATAN, R_LT, R_LE, R_EQ, R_NE, R_NOT, R_AND, R_OR, R_FLOOR, R_CEIL,
  R_FMOD, R_FABS, R_SIGN, R_COPYSIGN, R_IF_ELSE_ZERO, R_ERF, R_FMIN, R_FMAX, R_INV, R_SINH, R_COSH, R_TANH,
  R_ASINH, R_ACOSH, R_ATANH, R_ATAN2, R_CONST = 44, R_INPUT = 45, R_OUTPUT = 46, R_PARAMETER = 47, R_CALL = 48,
  R_ERFINV = 86, R_PRINTME = 87, R_LIFT = 88, R_LOG1P = 93, R_EXPM1 = 94, R_HYPOT = 95, R_REMAINDER = 97
};

// reference opcode -> (device opcode, number of slot operands); 0 operands = not evaluable
bool map_op(int rop, int* dop, int* nop) {
  switch (rop) {
    case R_ASSIGN: *dop = D_COPY; *nop = 1; return true;
    case R_ADD: *dop = D_ADD; *nop = 2; return true;
    case R_SUB: *dop = D_SUB; *nop = 2; return true;
    case R_MUL: *dop = D_MUL; *nop = 2; return true;
    case R_DIV: *dop = D_DIV; *nop = 2; return true;
    case R_NEG: *dop = D_NEG; *nop = 1; return true;
    case R_EXP: *dop = D_EXP; *nop = 1; return true;
    case R_LOG: *dop = D_LOG; *nop = 1; return true;
    case R_POW: *dop = D_POW; *nop = 2; return true;
    case R_CONSTPOW: *dop = D_POW; *nop = 2; return true;  // constpow(x,y) = pow(x,y), calculus.hpp:278
    case R_SQRT: *dop = D_SQRT; *nop = 1; return true;
    case R_SQ: *dop = D_SQ; *nop = 1; return true;
    case R_TWICE: *dop = D_TWICE; *nop = 1; return true;
    case R_SIN: *dop = D_SIN; *nop = 1; return true;
    case R_COS: *dop = D_COS; *nop = 1; return true;
    case R_TAN: *dop = D_TAN; *nop = 1; return true;
    case R_ASIN: *dop = D_ASIN; *nop = 1; return true;
    case R_ACOS: *dop = D_ACOS; *nop = 1; return true;
    case R_ATAN: *dop = D_ATAN; *nop = 1; return true;
    case R_LT: *dop = D_LT; *nop = 2; return true;
    case R_LE: *dop = D_LE; *nop = 2; return true;
    case R_EQ: *dop = D_EQ; *nop = 2; return true;
    case R_NE: *dop = D_NE; *nop = 2; return true;
    case R_NOT: *dop = D_NOT; *nop = 1; return true;
    case R_AND: *dop = D_AND; *nop = 2; return true;
    case R_OR: *dop = D_OR; *nop = 2; return true;
    case R_FLOOR: *dop = D_FLOOR; *nop = 1; return true;
    case R_CEIL: *dop = D_CEIL; *nop = 1; return true;
    case R_FMOD: *dop = D_FMOD; *nop = 2; return true;
    case R_FABS: *dop = D_FABS; *nop = 1; return true;
    case R_SIGN: *dop = D_SIGN; *nop = 1; return true;
    case R_COPYSIGN: *dop = D_COPYSIGN; *nop = 2; return true;
    case R_IF_ELSE_ZERO: *dop = D_IF_ELSE_ZERO; *nop = 2; return true;
    case R_ERF: *dop = D_ERF; *nop = 1; return true;
    case R_FMIN: *dop = D_FMIN; *nop = 2; return true;
    case R_FMAX: *dop = D_FMAX; *nop = 2; return true;
    case R_INV: *dop = D_INV; *nop = 1; return true;
    case R_SINH: *dop = D_SINH; *nop = 1; return true;
    case R_COSH: *dop = D_COSH; *nop = 1; return true;
    case R_TANH: *dop = D_TANH; *nop = 1; return true;
    case R_ASINH: *dop = D_ASINH; *nop = 1; return true;
    case R_ACOSH: *dop = D_ACOSH; *nop = 1; return true;
    case R_ATANH: *dop = D_ATANH; *nop = 1; return true;
    case R_ATAN2: *dop = D_ATAN2; *nop = 2; return true;
    case R_ERFINV: *dop = D_ERFINV; *nop = 1; return true;
    case R_LIFT: *dop = D_COPY; *nop = 1; return true;  // lift(x,y) = x, calculus.hpp:1007
    case R_LOG1P: *dop = D_LOG1P; *nop = 1; return true;
    case R_EXPM1: *dop = D_EXPM1; *nop = 1; return true;
    case R_HYPOT: *dop = D_HYPOT; *nop = 2; return true;
    case R_REMAINDER: *dop = D_REMAINDER; *nop = 2; return true;
    default: return false;
  }
}

}  // namespace

bool build_graph(const TapeSource& s, std::vector<Node>* nodes, long long* flops, std::string* err, long long* removed) {
  const long long n = s.n_instr;
  if (n < 0 || s.sz_w < 0) { *err = "negative tape size"; return false; }
  if (n > 0 && (!s.op || !s.i0 || !s.i1 || !s.i2 || !s.d)) { *err = "null tape arrays"; return false; }
  if (n >= (1ll << 31) - 2) { *err = "tape too long"; return false; }
  std::vector<int> def(static_cast<size_t>(s.sz_w) + 1, -1);  // slot -> defining node
  nodes->clear();
  nodes->reserve(n);
  *flops = 0;
  auto slot_ok = [&](int v) { return v >= 0 && v < s.sz_w; };
  for (long long k = 0; k < n; ++k) {
    Node nd;
    const int op = s.op[k];
    char buf[160];
    if (op == R_CONST) {
      if (!slot_ok(s.i0[k])) { snprintf(buf, sizeof buf, "instr %lld: slot out of range", k); *err = buf; return false; }
      nd.kind = K_CONST; nd.dop = D_CONST; nd.c = s.d[k];
      def[s.i0[k]] = static_cast<int>(k);
    } else if (op == R_INPUT) {
      if (!slot_ok(s.i0[k]) || s.i1[k] < 0 || s.i1[k] >= static_cast<int>(s.nnz_in.size()) || s.i2[k] < 0 ||
          s.i2[k] >= s.nnz_in[s.i1[k]]) {
        snprintf(buf, sizeof buf, "instr %lld: OP_INPUT index out of range", k); *err = buf; return false;
      }
      if (s.i1[k] > static_cast<int>(kMaxFieldIndex) || s.i2[k] > static_cast<int>(kMaxFieldIndex)) {
        snprintf(buf, sizeof buf, "instr %lld: input index exceeds device limit %u", k, kMaxFieldIndex);
        *err = buf; return false;
      }
      nd.kind = K_INPUT; nd.dop = D_INPUT; nd.idx = s.i1[k]; nd.nz = s.i2[k];
      def[s.i0[k]] = static_cast<int>(k);
    } else if (op == R_OUTPUT) {
      if (!slot_ok(s.i1[k]) || s.i0[k] < 0 || s.i0[k] >= static_cast<int>(s.nnz_out.size()) || s.i2[k] < 0 ||
          s.i2[k] >= s.nnz_out[s.i0[k]]) {
        snprintf(buf, sizeof buf, "instr %lld: OP_OUTPUT index out of range", k); *err = buf; return false;
      }
      if (s.i0[k] >= static_cast<int>(D_NONE) || s.i2[k] > static_cast<int>(kMaxFieldIndex)) {
        snprintf(buf, sizeof buf, "instr %lld: output index exceeds device limit", k); *err = buf; return false;
      }
      if (def[s.i1[k]] < 0) { snprintf(buf, sizeof buf, "instr %lld: reads undefined slot", k); *err = buf; return false; }
      nd.kind = K_OUTPUT; nd.dop = D_OUTPUT; nd.a = def[s.i1[k]]; nd.idx = s.i0[k]; nd.nz = s.i2[k];
    } else {
      int dop = 0, nop = 0;
      if (!map_op(op, &dop, &nop)) {
        const char* why = op == R_CALL ? "OP_CALL (embedded function call; expand() the function first)"
                        : op == R_PARAMETER ? "OP_PARAMETER (free variables cannot be evaluated, sx_function.cpp:78-83)"
                        : op == R_PRINTME ? "OP_PRINTME (host side effect)" : "not a scalar-evaluable operation";
        snprintf(buf, sizeof buf, "instr %lld: opcode %d unsupported on device: %s", k, op, why);
        *err = buf; return false;
      }
      if (!slot_ok(s.i0[k]) || !slot_ok(s.i1[k]) || (nop == 2 && !slot_ok(s.i2[k]))) {
        snprintf(buf, sizeof buf, "instr %lld: slot out of range", k); *err = buf; return false;
      }
      nd.kind = K_ARITH; nd.dop = static_cast<uint8_t>(dop);
      nd.a = def[s.i1[k]];
      nd.b = nop == 2 ? def[s.i2[k]] : -1;
      if (nd.a < 0 || (nop == 2 && nd.b < 0)) {
        snprintf(buf, sizeof buf, "instr %lld: reads undefined slot", k); *err = buf; return false;
      }
      def[s.i0[k]] = static_cast<int>(k);
      if (op != R_ASSIGN && op != R_LIFT) ++*flops;
    }
    nodes->push_back(nd);
  }
  // Value numbering: an instruction whose operation and operand VALUES equal those of an earlier one produces the same
  // bits (every operation is a pure function of its operands), so it is dropped and its readers use the earlier value.
  // The reference's AD emits the derivative's sin/cos, quotients and products next to identical primal ones (quadrotor
  // Jacobian: 9 100 of 76 911 instructions, among them 480 of its 960 sin/cos; rocket hess_lag: 4 596 of 20 321 incl.
  // 1 198 divisions).  ADD and MUL are matched in either operand order (IEEE addition and multiplication commute bit for
  // bit); fmin/fmax are not (their +-0 ties return the first operand).  `flops` stays the reference's count.
  static const bool cse_on = [] { const char* e = getenv("CCU_CSE"); return !(e && e[0] == '0'); }();
  if (removed) *removed = 0;
  if (cse_on && !nodes->empty()) {
    struct Key {
      uint64_t k0, k1;
      bool operator==(const Key& o) const { return k0 == o.k0 && k1 == o.k1; }
    };
    struct KeyHash {
      size_t operator()(const Key& k) const { return static_cast<size_t>((k.k0 * 0x9e3779b97f4a7c15ull) ^ (k.k1 + 0x7f4a7c15ull + (k.k0 << 6))); }
    };
    std::unordered_map<Key, int, KeyHash> seen;
    seen.reserve(nodes->size() * 2);
    const int m = static_cast<int>(nodes->size());
    std::vector<int> rep(m);      // node -> representative (old index)
    std::vector<int> newid(m, -1);
    std::vector<Node> kept;
    kept.reserve(m);
    long long dropped = 0;
    for (int k = 0; k < m; ++k) {
      Node nd = (*nodes)[k];
      rep[k] = k;
      if (nd.a >= 0) nd.a = newid[rep[nd.a]];
      if (nd.b >= 0) nd.b = newid[rep[nd.b]];
      if (nd.kind == K_OUTPUT) { newid[k] = static_cast<int>(kept.size()); kept.push_back(nd); continue; }
      Key key;
      if (nd.kind == K_CONST) {
        uint64_t bits;
        std::memcpy(&bits, &nd.c, 8);
        key = {(1ull << 62) | 1, bits};
      } else if (nd.kind == K_INPUT) {
        key = {(1ull << 62) | 2, (static_cast<uint64_t>(static_cast<uint32_t>(nd.idx)) << 32) | static_cast<uint32_t>(nd.nz)};
      } else {
        int a = nd.a, b = nd.b;
        if ((nd.dop == D_ADD || nd.dop == D_MUL) && b >= 0 && b < a) std::swap(a, b);
        key = {static_cast<uint64_t>(nd.dop), (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b)};
      }
      auto it = seen.find(key);
      if (it != seen.end()) {
        rep[k] = it->second;
        dropped += nd.kind == K_ARITH;
        continue;
      }
      seen.emplace(key, k);
      newid[k] = static_cast<int>(kept.size());
      kept.push_back(nd);
    }
    nodes->swap(kept);
    if (removed) *removed = dropped;
  }
  return true;
}

bool analyse_tape(const TapeSource& src, long long* max_live, long long* flops, std::string* err, long long* removed) {
  std::vector<Node> nodes;
  if (!build_graph(src, &nodes, flops, err, removed)) return false;
  const int n = static_cast<int>(nodes.size());
  std::vector<int> last(n);
  for (int k = 0; k < n; ++k) last[k] = k;
  for (int k = 0; k < n; ++k) {
    if (nodes[k].a >= 0) last[nodes[k].a] = k;
    if (nodes[k].b >= 0) last[nodes[k].b] = k;
  }
  std::vector<int> ev(n + 2, 0);
  for (int k = 0; k < n; ++k)
    if (nodes[k].kind != K_OUTPUT && last[k] > k) { ev[k] += 1; ev[last[k]] -= 1; }
  long long live = 0, mx = 0;
  for (int k = 0; k <= n; ++k) { live += ev[k]; mx = std::max(mx, live + 1); }
  *max_live = mx;
  return true;
}

bool compile_tape(const TapeSource& src, const CompileOptions& opt, Program* out, std::string* err) {
  std::vector<Node> nodes;
  long long flops = 0;
  if (!build_graph(src, &nodes, &flops, err)) return false;
  if (opt.schedule == 1) {
    // any topological order computes the same bits; this one keeps the live set small (tape_schedule.hpp)
    ScheduleOptions so;
    so.method = 1;
    so.seg_instr = 1 << 30;
    Schedule sch;
    if (!schedule_tape(nodes, so, &sch, err)) return false;
    std::vector<Node> ordered;
    permute_nodes(nodes, sch.order, &ordered);
    nodes.swap(ordered);
  }
  const int n = static_cast<int>(nodes.size());
  const int S = opt.slots_shared;
  if (S < 4) { *err = "slots_shared must be >= 4"; return false; }
  if (S > static_cast<int>(kMaxSharedSlots)) { *err = "slots_shared exceeds device limit"; return false; }

  // use lists (positions = node indices, increasing)
  std::vector<int> nuse(n, 0);
  for (int k = 0; k < n; ++k) {
    if (nodes[k].a >= 0) nuse[nodes[k].a]++;
    if (nodes[k].b >= 0 && nodes[k].b != nodes[k].a) nuse[nodes[k].b]++;
  }
  std::vector<int> ustart(n + 1, 0);
  for (int k = 0; k < n; ++k) ustart[k + 1] = ustart[k] + nuse[k];
  std::vector<int> uses(ustart[n]);
  {
    std::vector<int> fill(ustart.begin(), ustart.end() - 1);
    for (int k = 0; k < n; ++k) {
      if (nodes[k].a >= 0) uses[fill[nodes[k].a]++] = k;
      if (nodes[k].b >= 0 && nodes[k].b != nodes[k].a) uses[fill[nodes[k].b]++] = k;
    }
  }
  std::vector<int> uptr(ustart.begin(), ustart.end() - 1);  // next unread use of each value
  const int INF = 0x7fffffff;
  auto next_use = [&](int v) { return uptr[v] < ustart[v + 1] ? uses[uptr[v]] : INF; };

  std::vector<int> sslot(n, -1);  // value -> shared slot (or -1)
  std::vector<int> gslot(n, -1);  // value -> scratch slot holding a copy (or -1)
  std::vector<int> free_s, free_g;
  for (int i = S - 1; i >= 0; --i) free_s.push_back(i);
  int g_high = 0, s_high = 0;
  std::set<std::pair<int, int>> resident;  // (next use, value) of values in shared slots
  int acc_value = -1;                      // value currently in the forwarding register

  Program& P = *out;
  P = Program();
  P.n_instr = n;
  P.flops = flops;
  P.words.reserve(static_cast<size_t>(n) + 16);
  long long live = 0;

  auto emit = [&](uint32_t op, uint32_t d, uint32_t a, uint32_t b) { P.words.push_back(enc(op, d, a, b)); };
  auto alloc_g = [&]() {
    if (!free_g.empty()) { int g = free_g.back(); free_g.pop_back(); return g; }
    return g_high++;
  };
  // obtain a free shared slot, evicting the resident value with the furthest next use
  // (never one of the values in `pin`).  Evicted arithmetic values are spilled once.
  auto take_slot = [&](int pin0, int pin1) -> int {
    if (!free_s.empty()) { int s = free_s.back(); free_s.pop_back(); s_high = std::max(s_high, s + 1); return s; }
    auto it = resident.end();
    while (it != resident.begin()) {
      --it;
      int v = it->second;
      if (v == pin0 || v == pin1) continue;
      int s = sslot[v];
      if (nodes[v].kind == K_ARITH && gslot[v] < 0) {
        gslot[v] = alloc_g();
        emit(D_SPILL, 0, static_cast<uint32_t>(gslot[v]), v == acc_value ? F_ACC : static_cast<uint32_t>(s));
        P.spill_stores++;
        if (v != acc_value) P.smem_loads++;
      }
      sslot[v] = -1;
      resident.erase(it);
      return s;
    }
    return -1;
  };
  // make value v readable from a shared slot (or the forwarding register); returns operand field
  auto materialise = [&](int v, int pin_other) -> int {
    if (sslot[v] >= 0) return sslot[v];
    int s = take_slot(v, pin_other);
    if (s < 0) return -1;
    const Node& nv = nodes[v];
    if (nv.kind == K_CONST) {
      emit(D_CONST, static_cast<uint32_t>(s), 0, 0);
      uint64_t bits; std::memcpy(&bits, &nv.c, 8); P.words.push_back(bits);
      P.remat++;
    } else if (nv.kind == K_INPUT) {
      emit(D_INPUT, static_cast<uint32_t>(s), static_cast<uint32_t>(nv.idx), static_cast<uint32_t>(nv.nz));
      P.remat++;
    } else {
      emit(D_FILL, static_cast<uint32_t>(s), static_cast<uint32_t>(gslot[v]), 0);
      P.spill_loads++;
    }
    P.smem_stores++;
    acc_value = v;
    sslot[v] = s;
    resident.insert({next_use(v), v});
    return s;
  };
  // after instruction k has read value v: advance its use pointer, release it when dead
  auto consumed = [&](int v, int k) {
    if (sslot[v] >= 0) resident.erase({next_use(v), v});
    while (uptr[v] < ustart[v + 1] && uses[uptr[v]] <= k) ++uptr[v];
    if (next_use(v) == INF) {
      if (sslot[v] >= 0) { free_s.push_back(sslot[v]); sslot[v] = -1; }
      if (gslot[v] >= 0) { free_g.push_back(gslot[v]); gslot[v] = -1; }
      --live;
    } else if (sslot[v] >= 0) {
      resident.insert({next_use(v), v});
    }
  };

  for (int k = 0; k < n; ++k) {
    const Node& nd = nodes[k];
    uint32_t fa = 0, fb = 0;
    // ---- operands -------------------------------------------------------------------------
    if (nd.a >= 0) {
      const int va = nd.a, vb = nd.b;
      const bool need_a = !(opt.use_acc && va == acc_value) && sslot[va] < 0;
      const bool need_b = vb >= 0 && vb != va && !(opt.use_acc && vb == acc_value) && sslot[vb] < 0;
      if (need_a || need_b) {
        // a FILL / re-materialisation overwrites the forwarding register: an operand that lives
        // only there is parked in a shared slot first
        const int ops[2] = {va, vb};
        for (int v : ops) {
          if (v >= 0 && opt.use_acc && v == acc_value && sslot[v] < 0) {
            int s = take_slot(va, vb);
            if (s < 0) { *err = "internal: no shared slot to park operand"; return false; }
            emit(D_COPY, static_cast<uint32_t>(s), F_ACC, 0);
            P.smem_stores++;
            sslot[v] = s;
            resident.insert({next_use(v), v});
          }
        }
        if (need_a && materialise(va, vb) < 0) { *err = "internal: no shared slot for operand"; return false; }
        if (need_b && materialise(vb, va) < 0) { *err = "internal: no shared slot for operand"; return false; }
      }
      bool a_acc, b_acc;
      a_acc = opt.use_acc && va == acc_value;
      b_acc = opt.use_acc && vb >= 0 && vb == acc_value;
      if (!a_acc && sslot[va] < 0) { *err = "internal: operand a lost"; return false; }
      if (vb >= 0 && !b_acc && sslot[vb] < 0) { *err = "internal: operand b lost"; return false; }
      fa = a_acc ? F_ACC : static_cast<uint32_t>(sslot[va]);
      if (vb >= 0) fb = b_acc ? F_ACC : static_cast<uint32_t>(sslot[vb]);
      if (!a_acc) P.smem_loads++;
      if (vb >= 0 && !b_acc) P.smem_loads++;
    }
    // ---- the instruction ------------------------------------------------------------------
    if (nd.kind == K_OUTPUT) {
      emit(D_OUTPUT, static_cast<uint32_t>(nd.idx), fa, static_cast<uint32_t>(nd.nz));
      consumed(nd.a, k);
      continue;
    }
    if (nd.a >= 0) consumed(nd.a, k);
    if (nd.b >= 0 && nd.b != nd.a) consumed(nd.b, k);
    const int nu = next_use(k);
    uint32_t fd;
    bool spill_now = false;
    if (nu == INF) {
      fd = D_NONE;  // dead value (never read): compute and drop
    } else {
      ++live;
      P.max_live = std::max(P.max_live, live);
      // forward through the register when the only reader is the next instruction and that
      // instruction needs no FILL (so nothing overwrites the register in between)
      bool fwd = false;
      if (opt.use_acc && ustart[k + 1] - ustart[k] == 1 && nu == k + 1) {
        const Node& nx = nodes[k + 1];
        int other = nx.a == k ? nx.b : nx.a;
        fwd = other < 0 || other == k || sslot[other] >= 0;
      }
      if (fwd) {
        fd = D_NONE;
      } else {
        int s = -1;
        if (free_s.empty() && !resident.empty() && std::prev(resident.end())->first < nu) {
          // every resident value is needed sooner than this one: arithmetic results go straight to
          // the scratch; constants and inputs are simply re-materialised at their first use
          if (nd.kind != K_ARITH) continue;
          spill_now = true;
          fd = D_NONE;
        } else {
          s = take_slot(-1, -1);
          if (s < 0) { *err = "internal: no shared slot for result"; return false; }
          fd = static_cast<uint32_t>(s);
          sslot[k] = s;
          resident.insert({nu, k});
          P.smem_stores++;
        }
      }
    }
    if (nd.kind == K_CONST) {
      emit(D_CONST, fd, 0, 0);
      uint64_t bits; std::memcpy(&bits, &nd.c, 8); P.words.push_back(bits);
    } else if (nd.kind == K_INPUT) {
      emit(D_INPUT, fd, static_cast<uint32_t>(nd.idx), static_cast<uint32_t>(nd.nz));
    } else {
      emit(nd.dop, fd, fa, fb);
    }
    acc_value = k;
    if (spill_now) {
      gslot[k] = alloc_g();
      emit(D_SPILL, 0, static_cast<uint32_t>(gslot[k]), F_ACC);
      P.spill_stores++;
    }
  }
  emit(D_END, 0, 0, 0);
  P.slots_shared = std::max(s_high, 1);
  P.slots_global = g_high;
  if (g_high > static_cast<int>(kMaxFieldIndex)) { *err = "scratch slots exceed device limit"; return false; }
  return true;
}

}  // namespace ccu
