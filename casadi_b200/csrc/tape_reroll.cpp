// Loop re-rolling of an unrolled time-stepping tape.  See tape_reroll.hpp.
#include "tape_reroll.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <utility>

#include "ccu_isa.h"

namespace ccu {
namespace {

inline uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 29);
}
inline bool commutative(uint8_t dop) { return dop == D_ADD || dop == D_MUL; }  // IEEE +,* commute bit for bit

struct Finder {
  const std::vector<Node>& N;
  const int n;
  std::vector<int> depth;
  std::vector<uint64_t> h;
  std::vector<int> pi;      // arithmetic node -> its counterpart one iteration earlier (-1 unknown, itself = invariant)
  std::vector<int> tpi;     // tentative pairs of the running trial
  std::vector<int> touched;

  explicit Finder(const std::vector<Node>& nodes) : N(nodes), n(static_cast<int>(nodes.size())) {}

  bool arith(int v) const { return N[v].kind == K_ARITH; }
  // leaves pair with leaves of the same sort: constants with constants (their values may differ between iterations),
  // inputs with nonzeros of the same input (the nonzero may advance with the iteration)
  bool leaf_ok(int x, int y) const {
    if (x == y) return true;
    if (N[x].kind != N[y].kind) return false;
    if (N[x].kind == K_CONST) return true;
    if (N[x].kind == K_INPUT) return N[x].idx == N[y].idx;
    return false;
  }
  bool compat(int x, int y) const {
    if (arith(x) != arith(y)) return false;
    if (!arith(x)) return leaf_ok(x, y);
    if (pi[x] >= 0) return pi[x] == y;
    if (tpi[x] >= 0) return tpi[x] == y;
    return N[x].dop == N[y].dop;
  }
  // operand pairing of v -> w (swapped for a commutative operation when only that order is compatible)
  bool operand_pairs(int v, int w, int out[2][2]) const {
    const Node &a = N[v], &b = N[w];
    out[0][0] = a.a; out[0][1] = b.a; out[1][0] = a.b; out[1][1] = b.b;
    if (a.b < 0 || b.b < 0) return (a.b < 0) == (b.b < 0) && compat(a.a, b.a);
    const bool direct = compat(a.a, b.a) && compat(a.b, b.b);
    if (direct) {
      if (!commutative(a.dop)) return true;
      // both orders possible: prefer the one whose shallow hashes agree
      const bool swapped = compat(a.a, b.b) && compat(a.b, b.a);
      if (!swapped) return true;
      const bool hd = h[a.a] == h[b.a] && h[a.b] == h[b.b], hs = h[a.a] == h[b.b] && h[a.b] == h[b.a];
      if (hd || !hs) return true;
      out[0][1] = b.b; out[1][1] = b.a;
      return true;
    }
    if (commutative(a.dop) && compat(a.a, b.b) && compat(a.b, b.a)) {
      out[0][1] = b.b; out[1][1] = b.a;
      return true;
    }
    return false;
  }
  // does pairing v0 -> w0 hold for everything above `min_depth`?  On success the new pairs are left in tpi / touched.
  bool trial(int v0, int w0, int min_depth) {
    for (int v : touched) tpi[v] = -1;
    touched.clear();
    std::vector<std::pair<int, int>> st{{v0, w0}};
    while (!st.empty()) {
      const int v = st.back().first, w = st.back().second;
      st.pop_back();
      if (!arith(v) || !arith(w)) {
        if (arith(v) != arith(w) || !leaf_ok(v, w)) return false;
        continue;
      }
      if (N[v].dop != N[w].dop) return false;
      const int cur = pi[v] >= 0 ? pi[v] : tpi[v];
      if (cur >= 0) {
        if (cur != w) return false;
        continue;
      }
      tpi[v] = w;
      touched.push_back(v);
      if (v == w || depth[v] < min_depth) continue;
      int pr[2][2];
      if (!operand_pairs(v, w, pr)) return false;
      st.push_back({pr[0][0], pr[0][1]});
      if (pr[1][0] >= 0) st.push_back({pr[1][0], pr[1][1]});
    }
    return true;
  }
  void commit_trial() {
    for (int v : touched) { pi[v] = tpi[v]; tpi[v] = -1; }
    touched.clear();
  }
  // close pi over the operands of every paired node (no depth limit; a branch that does not match is left unpaired)
  void close_down() {
    std::vector<int> st;
    for (int v = 0; v < n; ++v) if (pi[v] >= 0 && pi[v] != v) st.push_back(v);
    while (!st.empty()) {
      const int v = st.back();
      st.pop_back();
      const int w = pi[v];
      if (!arith(v) || !arith(w) || N[v].dop != N[w].dop || v == w) continue;
      int pr[2][2];
      if (!operand_pairs(v, w, pr)) continue;
      for (int s = 0; s < 2; ++s) {
        const int x = pr[s][0], y = pr[s][1];
        if (x < 0 || !arith(x) || !arith(y) || pi[x] >= 0) continue;
        pi[x] = y;
        st.push_back(x);
      }
    }
  }
};

struct Key {
  uint64_t k0, k1;
  bool operator==(const Key& o) const { return k0 == o.k0 && k1 == o.k1; }
};
struct KeyHash {
  size_t operator()(const Key& k) const { return static_cast<size_t>(mix(k.k0, k.k1)); }
};
Key key_of(uint8_t dop, int a, int b) {
  if (commutative(dop) && b >= 0 && b < a) std::swap(a, b);
  return Key{static_cast<uint64_t>(dop), (static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b)};
}

}  // namespace

bool find_loop(const std::vector<Node>& N, Roll* R, int min_iters) {
  *R = Roll();
  Finder F(N);
  const int n = F.n;
  if (n < 48) { R->why = "tape too short"; return false; }
  // ---- depth and a shallow structural hash (value-independent for constants, nonzero-independent for inputs)
  F.depth.assign(n, 0);
  F.h.assign(n, 0);
  for (int v = 0; v < n; ++v) {
    const Node& nd = N[v];
    if (nd.kind == K_ARITH) {
      F.depth[v] = 1 + std::max(F.depth[nd.a], nd.b >= 0 ? F.depth[nd.b] : 0);
      F.h[v] = mix(19, nd.dop);
    } else if (nd.kind == K_CONST) {
      F.h[v] = 11;
    } else if (nd.kind == K_INPUT) {
      F.h[v] = mix(13, static_cast<uint64_t>(nd.idx));
    } else {
      F.h[v] = 17;
    }
  }
  {
    std::vector<uint64_t> h2(n);
    for (int round = 0; round < 6; ++round) {
      for (int v = 0; v < n; ++v) {
        const Node& nd = N[v];
        if (nd.kind != K_ARITH) { h2[v] = F.h[v]; continue; }
        const uint64_t ha = F.h[nd.a], hb = nd.b >= 0 ? F.h[nd.b] : 0x777;
        h2[v] = mix(mix(1000 + nd.dop, commutative(nd.dop) ? ha + hb : mix(ha, hb ^ 0x5555)), 1);
      }
      F.h.swap(h2);
    }
  }
  std::unordered_map<uint64_t, std::vector<int>> cls;
  for (int v = 0; v < n; ++v) if (F.arith(v)) cls[F.h[v]].push_back(v);
  // ---- the states the orbit starts from: the values the outputs read
  std::vector<int> outs;
  {
    std::vector<char> seen(n, 0);
    for (int v = 0; v < n; ++v)
      if (N[v].kind == K_OUTPUT && F.arith(N[v].a) && !seen[N[v].a]) { seen[N[v].a] = 1; outs.push_back(N[v].a); }
  }
  if (outs.empty()) { R->why = "no computed output"; return false; }
  std::sort(outs.begin(), outs.end(), [&](int x, int y) { return F.depth[x] != F.depth[y] ? F.depth[x] > F.depth[y] : x < y; });
  F.pi.assign(n, -1);
  F.tpi.assign(n, -1);
  // ---- anchors: pair every output value with its counterpart one iteration earlier
  {
    std::vector<int> stamp(n, -1), st, cand;
    int anchored = 0;
    for (size_t oi = 0; oi < outs.size(); ++oi) {
      const int o = outs[oi];
      if (F.pi[o] >= 0) { ++anchored; continue; }
      st.assign(1, o);
      stamp[o] = static_cast<int>(oi);
      while (!st.empty()) {  // ancestors of o
        const int v = st.back();
        st.pop_back();
        const int ops[2] = {N[v].a, N[v].b};
        for (int u : ops)
          if (u >= 0 && F.arith(u) && stamp[u] != static_cast<int>(oi)) { stamp[u] = static_cast<int>(oi); st.push_back(u); }
      }
      cand.clear();
      for (int w : cls[F.h[o]]) if (w != o && stamp[w] == static_cast<int>(oi)) cand.push_back(w);
      std::sort(cand.begin(), cand.end(), [&](int x, int y) { return F.depth[x] != F.depth[y] ? F.depth[x] > F.depth[y] : x > y; });
      if (cand.size() > 8) cand.resize(8);
      for (int w : cand) {
        const int period = F.depth[o] - F.depth[w];
        if (period <= 0) continue;
        if (F.trial(o, w, F.depth[o] - period)) { F.commit_trial(); ++anchored; break; }
      }
      for (int v : F.touched) F.tpi[v] = -1;
      F.touched.clear();
    }
    if (anchored == 0) { R->why = "no output has a counterpart one step earlier"; return false; }
  }
  F.close_down();
  // ---- upwards through the value-numbering table: a node whose operands all have counterparts has the counterpart
  //      (op, pi(a), pi(b)) when that node exists
  {
    std::unordered_map<Key, int, KeyHash> table;
    table.reserve(static_cast<size_t>(n) * 2);
    for (int v = 0; v < n; ++v) if (F.arith(v)) table.emplace(key_of(N[v].dop, N[v].a, N[v].b), v);
    for (int v = 0; v < n; ++v) {
      if (!F.arith(v) || F.pi[v] >= 0) continue;
      const Node& nd = N[v];
      if (!F.arith(nd.a) || F.pi[nd.a] < 0) continue;
      if (nd.b >= 0 && (!F.arith(nd.b) || F.pi[nd.b] < 0)) continue;
      auto it = table.find(key_of(nd.dop, F.pi[nd.a], nd.b >= 0 ? F.pi[nd.b] : -1));
      if (it != table.end()) F.pi[v] = it->second;
    }
    F.close_down();
  }
  auto valid = [&](int v) {
    const int w = F.pi[v];
    return w >= 0 && w != v && F.arith(v) && F.arith(w) && N[v].dop == N[w].dop;
  };
  // ---- orbit of the output values
  std::vector<std::vector<int>> S;
  S.push_back(outs);
  {
    std::vector<int> cur;
    for (int v : outs) if (valid(v)) cur.push_back(v);
    std::vector<char> used(n, 0);
    while (!cur.empty() && static_cast<int>(S.size()) < 100000) {
      std::vector<int> nxt;
      for (int v : cur) { const int w = F.pi[v]; if (!used[w]) { used[w] = 1; nxt.push_back(w); } }
      S.push_back(nxt);
      std::vector<int> c2;
      for (int v : nxt) if (valid(v)) c2.push_back(v);
      if (2 * c2.size() < nxt.size()) break;
      cur.swap(c2);
    }
  }
  const int KA = static_cast<int>(S.size());
  if (KA < min_iters + 1) { R->why = "orbit of the outputs is too short"; return false; }
  // ---- a node belongs to the first state that needs it (the deepest state first)
  std::vector<int> label(n, -1);
  {
    std::vector<int> st;
    for (int idx = KA - 1; idx >= 0; --idx) {
      for (int v : S[idx]) if (label[v] < 0) { label[v] = idx; st.push_back(v); }
      while (!st.empty()) {
        const int v = st.back();
        st.pop_back();
        const int ops[2] = {N[v].a, N[v].b};
        for (int u : ops)
          if (u >= 0 && F.arith(u) && label[u] < 0) { label[u] = idx; st.push_back(u); }
      }
    }
  }
  std::vector<std::vector<int>> body(KA);
  for (int v = 0; v < n; ++v) if (F.arith(v) && label[v] >= 0) body[label[v]].push_back(v);
  // ---- pi maps body idx onto body idx+1, one to one?
  std::vector<char> ok(KA, 0), hit(n, 0);
  for (int idx = 0; idx + 1 < KA; ++idx) {
    if (body[idx].empty() || body[idx].size() != body[idx + 1].size()) continue;
    bool good = true;
    for (int v : body[idx]) {
      if (!valid(v) || label[F.pi[v]] != idx + 1 || hit[F.pi[v]]) { good = false; break; }
      hit[F.pi[v]] = 1;
    }
    for (int v : body[idx]) if (F.pi[v] >= 0) hit[F.pi[v]] = 0;
    ok[idx] = good;
  }
  int best_lo = -1, best_len = 0;
  for (int idx = 0; idx + 1 < KA;) {
    if (!ok[idx]) { ++idx; continue; }
    int j = idx;
    while (j + 1 < KA && ok[j]) ++j;
    // labels idx .. j are iterations (ok[idx..j-1])
    if (j - idx + 1 > best_len) { best_len = j - idx + 1; best_lo = idx; }
    idx = j + 1;
  }
  if (best_len < min_iters) { R->why = "no run of identical steps"; return false; }
  const int lo = best_lo, hi = best_lo + best_len - 1, K = best_len;
  const int B = static_cast<int>(body[lo].size());
  R->iters = K;
  R->body = B;
  R->at.assign(K, std::vector<int>(B, -1));
  R->at[K - 1] = body[lo];  // ascending node ids: a topological order of the last iteration, hence of every iteration
  for (int t = K - 1; t > 0; --t)
    for (int p = 0; p < B; ++p) R->at[t - 1][p] = F.pi[R->at[t][p]];
  R->where.assign(n, -1);
  R->pos.assign(n, -1);
  for (int v = 0; v < n; ++v) {
    if (N[v].kind == K_OUTPUT) { R->where[v] = K; continue; }
    if (!F.arith(v)) continue;
    if (label[v] < 0) continue;            // not needed by any output: evaluated before the loop (checked by the template)
    if (label[v] < lo) R->where[v] = K;
    else if (label[v] > hi) R->where[v] = -1;
  }
  for (int t = 0; t < K; ++t)
    for (int p = 0; p < B; ++p) { R->where[R->at[t][p]] = t; R->pos[R->at[t][p]] = p; }
  R->found = true;
  return true;
}

bool build_loop_template(const std::vector<Node>& N, const Roll& R, LoopTemplate* T, std::string* why) {
  *T = LoopTemplate();
  if (!R.found) { *why = R.why; return false; }
  const int K = R.iters, B = R.body, n = static_cast<int>(N.size());
  char buf[200];
  auto fail = [&](const char* what, int t, int p) {
    snprintf(buf, sizeof buf, "%s (iteration %d, position %d)", what, t, p);
    *why = buf;
    return false;
  };
  T->K = K;
  T->B = B;
  T->dop.resize(B);
  T->a.assign(B, LoopOperand());
  T->b.assign(B, LoopOperand());
  std::vector<int> carried_index(B, -1);
  auto carried_of = [&](int q) {
    if (carried_index[q] < 0) { carried_index[q] = static_cast<int>(T->carried.size()); T->carried.push_back(q); T->entry.push_back(-1); }
    return carried_index[q];
  };
  for (int p = 0; p < B; ++p) {
    const int v1 = R.at[K - 1][p];
    T->dop[p] = N[v1].dop;
    for (int t = 0; t < K; ++t) {
      const Node& nd = N[R.at[t][p]];
      if (nd.kind != K_ARITH || nd.dop != T->dop[p]) return fail("operation differs between iterations", t, p);
      if ((nd.b < 0) != (N[v1].b < 0)) return fail("arity differs between iterations", t, p);
    }
    const int nops = N[v1].b >= 0 ? 2 : 1;
    // The operands of the LAST iteration define the roles; iteration t must fill the same roles (a commutative
    // operation may list its operands in the other order).
    struct Role { int kind, q, node; };  // 0 same iteration, 1 previous iteration, 2 fixed node, 5 constant, 6 input
    Role role[2];
    for (int s = 0; s < nops; ++s) {
      const int x = s == 0 ? N[v1].a : N[v1].b;
      if (N[x].kind == K_ARITH) {
        if (R.where[x] == K - 1) role[s] = {0, R.pos[x], x};
        else if (R.where[x] == K - 2) role[s] = {1, R.pos[x], x};
        else if (R.where[x] == -1) role[s] = {2, -1, x};
        else return fail("operand from more than one iteration back", K - 1, p);
      } else if (N[x].kind == K_CONST) {
        role[s] = {5, -1, x};
      } else if (N[x].kind == K_INPUT) {
        role[s] = {6, -1, x};
      } else {
        return fail("unexpected operand", K - 1, p);
      }
    }
    auto fills = [&](const Role& r, int y, int t) {
      switch (r.kind) {
        case 0: return N[y].kind == K_ARITH && R.where[y] == t && R.pos[y] == r.q;
        case 1:
          if (t >= 1) return N[y].kind == K_ARITH && R.where[y] == t - 1 && R.pos[y] == r.q;
          return N[y].kind == K_CONST || N[y].kind == K_INPUT || (N[y].kind == K_ARITH && R.where[y] == -1);
        case 2: return y == r.node;
        case 5: return N[y].kind == K_CONST;
        default: return N[y].kind == K_INPUT && N[y].idx == N[r.node].idx;
      }
    };
    std::vector<int> opnd[2];  // operand of every iteration, per role
    opnd[0].resize(K);
    opnd[1].resize(K);
    for (int t = 0; t < K; ++t) {
      const Node& nd = N[R.at[t][p]];
      if (nops == 1) {
        if (!fills(role[0], nd.a, t)) return fail("operand is irregular", t, p);
        opnd[0][t] = nd.a;
      } else if (fills(role[0], nd.a, t) && fills(role[1], nd.b, t)) {
        opnd[0][t] = nd.a; opnd[1][t] = nd.b;
      } else if (commutative(nd.dop) && fills(role[0], nd.b, t) && fills(role[1], nd.a, t)) {
        opnd[0][t] = nd.b; opnd[1][t] = nd.a;
      } else {
        return fail("operands are irregular", t, p);
      }
    }
    for (int s = 0; s < nops; ++s) {
      LoopOperand& out = s == 0 ? T->a[p] : T->b[p];
      const Role& r = role[s];
      if (r.kind == 0) {
        out.kind = 0;
        out.ref = r.q;
      } else if (r.kind == 1) {
        out.kind = 1;
        out.ref = r.q;
        const int ci = carried_of(r.q);
        const int y0 = opnd[s][0];
        if (T->entry[ci] >= 0 && T->entry[ci] != y0) {
          // two constants with the same bits are the same entry value
          const Node &e0 = N[T->entry[ci]], &e1 = N[y0];
          if (!(e0.kind == K_CONST && e1.kind == K_CONST && std::memcmp(&e0.c, &e1.c, 8) == 0)) return fail("entry value is ambiguous", 0, p);
        }
        T->entry[ci] = y0;
      } else if (r.kind == 2) {
        out.kind = 2;
        out.ref = r.node;
      } else if (r.kind == 5) {
        bool same = true;
        std::vector<double> col(K);
        for (int t = 0; t < K; ++t) {
          col[t] = N[opnd[s][t]].c;
          same = same && std::memcmp(&col[t], &N[r.node].c, 8) == 0;
        }
        if (same) { out.kind = 2; out.ref = r.node; }
        else { out.kind = 3; out.ref = static_cast<int>(T->ctab.size()); T->ctab.push_back(col); }
      } else {
        bool same = true;
        std::vector<int> nz(K);
        for (int t = 0; t < K; ++t) {
          nz[t] = N[opnd[s][t]].nz;
          same = same && nz[t] == N[r.node].nz;
        }
        if (same) { out.kind = 2; out.ref = r.node; }
        else {
          const int stride = nz[1] - nz[0];
          for (int t = 1; t < K; ++t) if (nz[t] - nz[t - 1] != stride) return fail("input nonzero does not advance evenly", t, p);
          out.kind = 4;
          out.ref = static_cast<int>(T->affine.size());
          T->affine.push_back({N[r.node].idx, nz[0], stride});
        }
      }
    }
  }
  // ---- uses from outside the loop: nothing before it reads it, whatever comes after reads the LAST iteration only
  std::vector<char> is_exit(B, 0);
  for (int v = 0; v < n; ++v) {
    const int wv = R.where[v];
    if (wv >= 0 && wv < K) continue;
    const int ops[2] = {N[v].kind == K_CONST || N[v].kind == K_INPUT ? -1 : N[v].a, N[v].kind == K_ARITH ? N[v].b : -1};
    for (int u : ops) {
      if (u < 0 || N[u].kind != K_ARITH) continue;
      const int wu = R.where[u];
      if (wu < 0 || wu >= K) {
        if (wv == -1 && wu == K) { *why = "a value computed before the loop depends on one computed after it"; return false; }
        continue;
      }
      if (wv == -1) { *why = "a value computed before the loop reads the loop"; return false; }
      if (wu != K - 1) { *why = "a value computed after the loop reads an earlier iteration"; return false; }
      is_exit[R.pos[u]] = 1;
    }
  }
  for (int p = 0; p < B; ++p) if (is_exit[p]) T->exit_pos.push_back(p);
  for (size_t c = 0; c < T->carried.size(); ++c)
    if (T->entry[c] < 0) { *why = "carried value without an entry"; return false; }
  return true;
}

}  // namespace ccu
