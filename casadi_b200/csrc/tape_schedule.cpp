// Min-cut recursive bisection of the tape's value graph.  See tape_schedule.hpp.
#include "tape_schedule.hpp"

#include "ccu_isa.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <set>
#include <utility>

namespace ccu {
namespace {

constexpr int kInf = 1 << 28;

struct PieceRec { int begin, end, arith; long long weight; };

// estimated SASS instructions of one tape instruction in a specialised kernel (sm_100a, --fmad=false; measured on the
// quadrotor tapes: division 14, sin/cos 45, everything with a libm slow path is long) -- a kernel has to stay within
// the instruction cache (a 167 KB primal segment of the Jacobian ran at 14 % issue utilisation, stall_no_instruction
// 9.2 per issue: profiles/r1_ncu_full_seg_jac_icache.txt)
int op_weight(const Node& nd) {
  if (nd.kind == K_OUTPUT) return 3;
  if (nd.kind != K_ARITH) return 0;
  switch (nd.dop) {
    case D_DIV: case D_INV: case D_SQRT: return 14;
    case D_SIN: case D_COS: return 45;
    case D_TAN: case D_EXP: case D_LOG: case D_LOG1P: case D_EXPM1: case D_SINH: case D_COSH: case D_TANH: return 50;
    case D_POW: case D_ATAN2: case D_HYPOT: case D_FMOD: case D_REMAINDER: case D_ASIN: case D_ACOS: case D_ATAN:
    case D_ASINH: case D_ACOSH: case D_ATANH: case D_ERF: case D_ERFINV: return 90;
    case D_COPY: return 0;
    default: return 2;
  }
}

// Dinic's maximum flow on a forward-star graph; blocking flows are found iteratively (dependency chains make
// augmenting paths thousands of edges long).
struct FlowNet {
  struct Edge { int to, cap; };
  std::vector<Edge> e;
  std::vector<int> head, nxt, level, it, queue, path;
  int n = 0;

  void reset(int nodes) {
    n = nodes;
    head.assign(n, -1);
    e.clear();
    nxt.clear();
  }
  int add_node() { head.push_back(-1); return n++; }
  void add(int a, int b, int c) {
    e.push_back({b, c}); nxt.push_back(head[a]); head[a] = static_cast<int>(e.size()) - 1;
    e.push_back({a, 0}); nxt.push_back(head[b]); head[b] = static_cast<int>(e.size()) - 1;
  }
  bool bfs(int s, int t) {
    level.assign(n, -1);
    queue.clear();
    queue.push_back(s);
    level[s] = 0;
    for (size_t q = 0; q < queue.size(); ++q) {
      const int u = queue[q];
      for (int ei = head[u]; ei != -1; ei = nxt[ei])
        if (e[ei].cap > 0 && level[e[ei].to] < 0) { level[e[ei].to] = level[u] + 1; queue.push_back(e[ei].to); }
    }
    return level[t] >= 0;
  }
  long long maxflow(int s, int t) {
    long long flow = 0;
    while (bfs(s, t)) {
      it = head;
      path.clear();
      int u = s;
      for (;;) {
        if (u == t) {
          int f = kInf;
          for (int ei : path) f = std::min(f, e[ei].cap);
          for (int ei : path) { e[ei].cap -= f; e[ei ^ 1].cap += f; }
          flow += f;
          size_t k = 0;
          while (e[path[k]].cap > 0) ++k;  // first saturated edge
          path.resize(k);
          u = k ? e[path[k - 1]].to : s;
          continue;
        }
        bool advanced = false;
        for (int& ei = it[u]; ei != -1; ei = nxt[ei]) {
          if (e[ei].cap > 0 && level[e[ei].to] == level[u] + 1) {
            path.push_back(ei);
            u = e[ei].to;
            advanced = true;
            break;
          }
        }
        if (advanced) continue;
        level[u] = -1;  // dead end
        if (path.empty()) break;
        const int ei = path.back();
        path.pop_back();
        u = e[ei ^ 1].to;
        it[u] = nxt[ei];
      }
    }
    return flow;
  }
  // nodes reachable from s in the residual graph
  void reach_from(int s, std::vector<char>* mark) {
    mark->assign(n, 0);
    queue.clear();
    queue.push_back(s);
    (*mark)[s] = 1;
    for (size_t q = 0; q < queue.size(); ++q) {
      const int u = queue[q];
      for (int ei = head[u]; ei != -1; ei = nxt[ei])
        if (e[ei].cap > 0 && !(*mark)[e[ei].to]) { (*mark)[e[ei].to] = 1; queue.push_back(e[ei].to); }
    }
  }
  // nodes that can reach t in the residual graph
  void reach_to(int t, std::vector<char>* mark) {
    mark->assign(n, 0);
    queue.clear();
    queue.push_back(t);
    (*mark)[t] = 1;
    for (size_t q = 0; q < queue.size(); ++q) {
      const int u = queue[q];
      for (int ei = head[u]; ei != -1; ei = nxt[ei])
        if (e[ei ^ 1].cap > 0 && !(*mark)[e[ei].to]) { (*mark)[e[ei].to] = 1; queue.push_back(e[ei].to); }
    }
  }
};

// The recursion is position based so that its two halves can run in parallel: a piece owns the interval
// [begin, begin + size) of the final order; `lo[v]` is the start of the smallest interval known to contain v.  For a
// piece [L, H) a node outside it lies entirely before L or at/after H, whatever other threads are refining meanwhile.
struct Bisector {
  const std::vector<Node>& N;
  const ScheduleOptions& opt;
  int n;
  std::vector<int> cstart, cons;      // distinct consumers of every arithmetic value
  std::vector<int> partner;           // sin(x) <-> cos(x) of the same operand (kept in one piece: jit.cpp fuses them)
  std::vector<std::atomic<int>> lo;   // node -> start of its current interval
  std::vector<int> out;               // items in final order (written by position)
  std::vector<PieceRec> pieces;       // every piece of the recursion (any order; sorted at the end)
  std::mutex pieces_mutex;
  std::atomic<long long> cuts{0};
  std::atomic<int> threads_free{0};

  // per-thread scratch
  struct Worker {
    std::vector<int> loc;    // node -> index in the current piece (-1 = outside)
    std::vector<int> ext_z;  // external value -> flow node of the current network (-1 = none, -2 = constant)
    std::vector<int> ext_touched;
    FlowNet net;
    explicit Worker(int n) : loc(n, -1), ext_z(n, -1) {}
  };

  Bisector(const std::vector<Node>& nodes, const ScheduleOptions& o)
      : N(nodes), opt(o), n(static_cast<int>(nodes.size())), lo(nodes.size()) {
    std::vector<int> cnt(n + 1, 0);
    auto each_operand = [&](int k, auto&& f) {
      const Node& nd = N[k];
      if (nd.a >= 0 && N[nd.a].kind == K_ARITH) f(nd.a);
      if (nd.b >= 0 && nd.b != nd.a && N[nd.b].kind == K_ARITH) f(nd.b);
    };
    for (int k = 0; k < n; ++k) each_operand(k, [&](int u) { cnt[u + 1]++; });
    cstart.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) cstart[k + 1] = cstart[k] + cnt[k + 1];
    cons.resize(cstart[n]);
    std::vector<int> fill(cstart.begin(), cstart.end() - 1);
    for (int k = 0; k < n; ++k) each_operand(k, [&](int u) { cons[fill[u]++] = k; });
    for (auto& x : lo) x.store(0, std::memory_order_relaxed);
    // sin / cos pairs on one operand: the specialised kernels evaluate a pair with ONE argument reduction when both
    // are in the same kernel (238 of the quadrotor Jacobian's 480 pairs were split by the cuts before they were tied)
    partner.assign(n, -1);
    if (opt.tie_sincos) {
      std::vector<int> first_sin(n, -1), first_cos(n, -1);
      for (int k = 0; k < n; ++k) {
        const Node& nd = N[k];
        if (nd.kind != K_ARITH || nd.a < 0) continue;
        if (nd.dop == D_SIN && first_sin[nd.a] < 0) first_sin[nd.a] = k;
        if (nd.dop == D_COS && first_cos[nd.a] < 0) first_cos[nd.a] = k;
      }
      for (int v = 0; v < n; ++v)
        if (first_sin[v] >= 0 && first_cos[v] >= 0) { partner[first_sin[v]] = first_cos[v]; partner[first_cos[v]] = first_sin[v]; }
    }
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* e = getenv("CCU_SCHED_THREADS")) hw = atoi(e);
    threads_free = std::max(0, std::min(hw, 16) - 1);
  }

  template <class F>
  void operands(int k, F&& f) const {
    const Node& nd = N[k];
    if (nd.a >= 0 && N[nd.a].kind == K_ARITH) f(nd.a);
    if (nd.b >= 0 && nd.b != nd.a && N[nd.b].kind == K_ARITH) f(nd.b);
  }

  // minimum cut of `piece` = interval [L, H) with the first/last npin items of `order` pinned; inD[i] for piece
  // index i.  Of the two extreme minimum cuts (smallest / largest D) the more balanced one is returned.
  long long mincut(Worker& W, const std::vector<int>& piece, const std::vector<int>& order, int npin, int H,
                   std::vector<char>* inD) {
    const int m = static_cast<int>(piece.size());
    const int S = 0, T = 1;
    FlowNet& net = W.net;
    std::vector<int>& loc = W.loc;
    std::vector<int>& ext_z = W.ext_z;
    net.reset(2 + m);
    auto X = [](int i) { return 2 + i; };
    for (int j = 0; j < npin; ++j) net.add(S, X(loc[order[j]]), kInf);
    for (int j = m - npin; j < m; ++j) net.add(X(loc[order[j]]), T, kInf);
    {  // tied pairs stay on one side of the cut (unless the pins already separate them)
      std::vector<char> pin(m, 0);
      for (int j = 0; j < npin; ++j) pin[loc[order[j]]] = 1;
      for (int j = m - npin; j < m; ++j) pin[loc[order[j]]] |= 2;
      for (int i = 0; i < m; ++i) {
        const int p = partner[piece[i]];
        if (p < 0 || loc[p] < 0 || loc[p] < i) continue;
        const int pi = loc[p];
        if (((pin[i] | pin[pi]) & 3) == 3) continue;
        net.add(X(i), X(pi), kInf);
        net.add(X(pi), X(i), kInf);
      }
    }
    W.ext_touched.clear();
    for (int i = 0; i < m; ++i) {
      const int v = piece[i];
      operands(v, [&](int u) {
        if (loc[u] >= 0) {
          net.add(X(i), X(loc[u]), kInf);  // closure: v in D => u in D
          return;
        }
        // value of an earlier piece: it stays live across this cut iff one of its readers is outside D.  When it
        // is also read after this piece it is live whatever the cut.
        if (ext_z[u] == -1) {
          bool later = false;
          for (int q = cstart[u]; q < cstart[u + 1]; ++q) {
            const int c = cons[q];
            if (loc[c] < 0 && lo[c].load(std::memory_order_relaxed) >= H) { later = true; break; }
          }
          W.ext_touched.push_back(u);
          if (later) {
            ext_z[u] = -2;
          } else {
            ext_z[u] = net.add_node();
            net.add(S, ext_z[u], 1);
          }
        }
        if (ext_z[u] >= 0) net.add(ext_z[u], X(i), kInf);
      });
      if (N[v].kind != K_ARITH) continue;
      const int nc = cstart[v + 1] - cstart[v];
      if (nc == 0) continue;
      bool outside = false;
      for (int q = cstart[v]; q < cstart[v + 1]; ++q)
        if (loc[cons[q]] < 0) { outside = true; break; }
      if (outside) {
        net.add(X(i), T, 1);  // read after this piece: live across the cut iff v is in D
      } else if (nc == 1) {
        net.add(X(i), X(loc[cons[cstart[v]]]), 1);
      } else {
        const int z = net.add_node();
        net.add(X(i), z, 1);
        for (int q = cstart[v]; q < cstart[v + 1]; ++q) net.add(z, X(loc[cons[q]]), kInf);
      }
    }
    for (int u : W.ext_touched) ext_z[u] = -1;
    const long long f = net.maxflow(S, T);
    ++cuts;
    std::vector<char> a, b;
    net.reach_from(S, &a);
    net.reach_to(T, &b);
    int na = 0, nb = 0;
    for (int i = 0; i < m; ++i) { na += a[X(i)]; nb += !b[X(i)]; }
    inD->assign(m, 0);
    const bool use_small = std::abs(2 * na - m) <= std::abs(2 * nb - m);
    for (int i = 0; i < m; ++i) (*inD)[i] = use_small ? a[X(i)] : !b[X(i)];
    return f;
  }

  int arith_count(const std::vector<int>& piece) const {
    int c = 0;
    for (int v : piece) c += N[v].kind == K_ARITH;
    return c;
  }

  void emit(const std::vector<int>& piece, int L) {
    for (size_t i = 0; i < piece.size(); ++i) out[L + i] = piece[i];
  }

  // piece occupies [L, L + piece.size()) of the final order
  void split(Worker& W, std::vector<int>& piece, int L) {
    const int m = static_cast<int>(piece.size());
    const int H = L + m;
    long long w = 0;
    for (int v : piece) w += op_weight(N[v]);
    {
      std::lock_guard<std::mutex> lock(pieces_mutex);
      pieces.push_back({L, H, arith_count(piece), w});
    }
    if (m <= std::max(opt.min_piece, 2)) {
      emit(piece, L);
      return;
    }
    std::vector<int>& loc = W.loc;
    for (int i = 0; i < m; ++i) loc[piece[i]] = i;
    const int npin = std::max(1, std::min(m / 2, static_cast<int>(opt.pin_frac * m)));
    // order 1: as inherited; order 2: ASAP levels inside the piece
    std::vector<int> lev(m, 0);
    int maxlev = 0;
    for (int i = 0; i < m; ++i) {
      int l = 0;
      operands(piece[i], [&](int u) { if (loc[u] >= 0) l = std::max(l, lev[loc[u]] + 1); });
      lev[i] = l;
      maxlev = std::max(maxlev, l);
    }
    std::vector<int> bucket(maxlev + 2, 0);
    for (int i = 0; i < m; ++i) bucket[lev[i] + 1]++;
    for (int l = 0; l <= maxlev; ++l) bucket[l + 1] += bucket[l];
    std::vector<int> by_level(m);
    for (int i = 0; i < m; ++i) by_level[bucket[lev[i]]++] = piece[i];
    std::vector<char> d1, d2;
    long long c1 = 0, c2 = kInf;
    bool both = false;
    if (maxlev > 0 && m > 4000) {  // the two cuts of a large piece side by side
      int f = threads_free.load();
      while (f > 0 && !threads_free.compare_exchange_weak(f, f - 1)) {}
      if (f > 0) {
        both = true;
        std::thread t2([&] {
          Worker W2(n);
          for (int i = 0; i < m; ++i) W2.loc[piece[i]] = i;
          c2 = mincut(W2, piece, by_level, npin, H, &d2);
          threads_free.fetch_add(1);
        });
        c1 = mincut(W, piece, piece, npin, H, &d1);
        t2.join();
      }
    }
    if (!both) {
      c1 = mincut(W, piece, piece, npin, H, &d1);
      if (maxlev > 0) c2 = mincut(W, piece, by_level, npin, H, &d2);
    }
    auto imbalance = [&](const std::vector<char>& d) {
      int k = 0;
      for (char x : d) k += x;
      return std::abs(2 * k - m);
    };
    const bool second = c2 < c1 || (c2 == c1 && imbalance(d2) < imbalance(d1));
    const std::vector<char>& d = second ? d2 : d1;
    for (int i = 0; i < m; ++i) loc[piece[i]] = -1;
    std::vector<int> A, B;
    for (int i = 0; i < m; ++i) (d[i] ? A : B).push_back(piece[i]);
    std::vector<int>().swap(piece);
    if (A.empty() || B.empty()) {  // cannot happen with pins on both sides; keep the order rather than loop
      emit(A.empty() ? B : A, L);
      return;
    }
    const int mid = L + static_cast<int>(A.size());
    for (int v : B) lo[v].store(mid, std::memory_order_relaxed);  // (A keeps L)
    // large halves go to another thread when one is free
    bool spawned = false;
    std::thread other;
    if (static_cast<int>(A.size()) > 2000 && static_cast<int>(B.size()) > 2000) {
      int f = threads_free.load();
      while (f > 0 && !threads_free.compare_exchange_weak(f, f - 1)) {}
      if (f > 0) {
        spawned = true;
        other = std::thread([this, &B, mid] {
          Worker W2(n);
          split(W2, B, mid);
          threads_free.fetch_add(1);
        });
      }
    }
    split(W, A, L);
    if (spawned) other.join();
    else split(W, B, mid);
  }
};

struct Recursion {
  int n = 0;
  std::vector<int> out;
  std::vector<PieceRec> pieces;
  long long cuts = 0;
};
std::mutex g_cache_mutex;
std::deque<std::pair<uint64_t, std::shared_ptr<const Recursion>>> g_cache;

}  // namespace

bool schedule_tape(const std::vector<Node>& nodes, const ScheduleOptions& opt, Schedule* S, std::string* err) {
  const auto t0 = std::chrono::steady_clock::now();
  const int n = static_cast<int>(nodes.size());
  *S = Schedule();
  const int per = std::max(opt.seg_instr, 16);
  // (the bisection costs seconds per 100 K instructions; very long tapes keep the reference order)
  if (opt.method == 0 || n == 0 || n > opt.max_nodes) {
    S->order.resize(n);
    for (int k = 0; k < n; ++k) S->order[k] = k;
    S->seg_begin.push_back(0);
    int cnt = 0;
    for (int k = 0; k < n; ++k)
      if (nodes[k].kind == K_ARITH && ++cnt >= per && k + 1 < n) { cnt = 0; S->seg_begin.push_back(k + 1); }
    S->seg_begin.push_back(n);
    return true;
  }
  // the recursion does not depend on seg_instr: it is computed once per tape and kept (a tape is planned several
  // times: interpreter order, specialised kernels, re-plans with other segment lengths)
  ScheduleOptions o = opt;
  o.min_piece = std::min(opt.min_piece, std::max(2, per / 2));  // (tiny segments are only asked for by tests)
  if (const char* e = getenv("CCU_SCHED_TIE")) o.tie_sincos = atoi(e);
  uint64_t key = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { key ^= v; key *= 1099511628211ull; };
  mix(static_cast<uint64_t>(n)); mix(static_cast<uint64_t>(o.min_piece)); mix(static_cast<uint64_t>(opt.pin_frac * 1e6));
  mix(static_cast<uint64_t>(o.tie_sincos));
  for (const Node& nd : nodes) {
    uint64_t cb;
    std::memcpy(&cb, &nd.c, 8);
    mix((static_cast<uint64_t>(nd.kind) << 8) | nd.dop); mix(static_cast<uint64_t>(static_cast<uint32_t>(nd.a)));
    mix(static_cast<uint64_t>(static_cast<uint32_t>(nd.b))); mix((static_cast<uint64_t>(nd.idx) << 32) | static_cast<uint32_t>(nd.nz));
    mix(cb);
  }
  std::shared_ptr<const Recursion> rec;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto& e : g_cache)
      if (e.first == key && e.second->n == n) { rec = e.second; break; }
  }
  if (!rec) {
    Bisector B(nodes, o);
    std::vector<int> items;
    items.reserve(n);
    for (int k = 0; k < n; ++k)
      if (nodes[k].kind == K_ARITH || nodes[k].kind == K_OUTPUT) items.push_back(k);
    const size_t n_items = items.size();
    B.out.assign(n_items, -1);
    if (!items.empty()) {
      Bisector::Worker W(n);
      B.split(W, items, 0);
    }
    for (int v : B.out)
      if (v < 0) { *err = "internal: schedule lost nodes"; return false; }
    // pre-order: by start, enclosing piece first
    std::sort(B.pieces.begin(), B.pieces.end(), [](const PieceRec& x, const PieceRec& y) {
      if (x.begin != y.begin) return x.begin < y.begin;
      return x.end > y.end;
    });
    auto r = std::make_shared<Recursion>();
    r->n = n;
    r->out = std::move(B.out);
    r->pieces = std::move(B.pieces);
    r->cuts = B.cuts.load();
    rec = r;
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cache.push_front({key, rec});
    if (g_cache.size() > 16) g_cache.pop_back();
  }
  const std::vector<int>& out = rec->out;
  // segments: the maximal pieces with at most `per` arithmetic instructions (pieces are in pre-order, so a piece
  // inside an accepted one starts before that one's end); adjacent small ones are merged
  std::vector<int> merged;
  {
    const long long wmax = opt.seg_weight > 0 ? opt.seg_weight : (1ll << 60);
    struct Seg { int begin, arith; long long weight; };
    std::vector<Seg> segs;
    int covered = 0;
    for (const PieceRec& pc : rec->pieces) {
      if (pc.begin < covered) continue;
      if ((pc.arith <= per && pc.weight <= wmax) || pc.end - pc.begin <= 2) { segs.push_back({pc.begin, pc.arith, pc.weight}); covered = pc.end; }
    }
    int cur = 0;
    long long curw = 0;
    for (const Seg& sgm : segs) {
      if (merged.empty() || cur + sgm.arith > per || curw + sgm.weight > wmax) { merged.push_back(sgm.begin); cur = sgm.arith; curw = sgm.weight; }
      else { cur += sgm.arith; curw += sgm.weight; }
    }
  }
  // constants and inputs go right before their first reader (they are re-materialised by both kernel families)
  std::vector<char> placed(n, 0);
  S->order.reserve(n);
  size_t ms = 0;
  for (size_t i = 0; i < out.size(); ++i) {
    if (ms < merged.size() && merged[ms] == static_cast<int>(i)) { S->seg_begin.push_back(static_cast<int>(S->order.size())); ++ms; }
    const int v = out[i];
    const int ops[2] = {nodes[v].a, nodes[v].b};
    for (int u : ops)
      if (u >= 0 && nodes[u].kind != K_ARITH && !placed[u]) { placed[u] = 1; S->order.push_back(u); }
    placed[v] = 1;
    S->order.push_back(v);
  }
  for (int k = 0; k < n; ++k)
    if (!placed[k]) S->order.push_back(k);  // unread constants / inputs
  if (S->seg_begin.empty()) S->seg_begin.push_back(0);
  S->seg_begin.push_back(n);
  // the order must be topological
  std::vector<int> pos(n, -1);
  if (static_cast<int>(S->order.size()) != n) { *err = "internal: schedule is not a permutation"; return false; }
  for (int i = 0; i < n; ++i) pos[S->order[i]] = i;
  for (int k = 0; k < n; ++k) {
    if (pos[k] < 0) { *err = "internal: schedule is not a permutation"; return false; }
    if ((nodes[k].a >= 0 && pos[nodes[k].a] >= pos[k]) || (nodes[k].b >= 0 && pos[nodes[k].b] >= pos[k])) {
      *err = "internal: schedule violates a dependency";
      return false;
    }
  }
  S->cuts = rec->cuts;
  S->ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

void permute_nodes(const std::vector<Node>& in, const std::vector<int>& order, std::vector<Node>* out) {
  const int n = static_cast<int>(in.size());
  std::vector<int> pos(n);
  for (int i = 0; i < n; ++i) pos[order[i]] = i;
  out->resize(n);
  for (int i = 0; i < n; ++i) {
    Node nd = in[order[i]];
    if (nd.a >= 0) nd.a = pos[nd.a];
    if (nd.b >= 0) nd.b = pos[nd.b];
    (*out)[i] = nd;
  }
}

void cross_traffic(const std::vector<Node>& nodes, const std::vector<int>& seg_begin, long long* loads, long long* stores) {
  const int n = static_cast<int>(nodes.size());
  std::vector<int> seg(n, 0);
  for (size_t s = 0; s + 1 < seg_begin.size(); ++s)
    for (int k = seg_begin[s]; k < seg_begin[s + 1]; ++k) seg[k] = static_cast<int>(s);
  std::set<std::pair<int, int>> ld;
  std::vector<char> st(n, 0);
  for (int k = 0; k < n; ++k) {
    const int ops[2] = {nodes[k].a, nodes[k].b};
    for (int u : ops)
      if (u >= 0 && nodes[u].kind == K_ARITH && seg[u] != seg[k]) { ld.insert({u, seg[k]}); st[u] = 1; }
  }
  *loads = static_cast<long long>(ld.size());
  *stores = 0;
  for (char c : st) *stores += c;
}

// ------------------------------------------------------------------------------------------ rematerialisation
namespace {
// FP64 issue slots of one recomputed instruction (branch-free division 9, sincos ~22 per pair, libm calls long)
int remat_cost(const Node& nd) {
  switch (nd.dop) {
    case D_COPY: return 0;
    case D_DIV: case D_INV: case D_SQRT: return 10;
    case D_SIN: case D_COS: return 14;
    case D_ADD: case D_SUB: case D_MUL: case D_NEG: case D_SQ: case D_TWICE: case D_FABS: case D_LT: case D_LE: case D_EQ:
    case D_NE: case D_NOT: case D_AND: case D_OR: case D_IF_ELSE_ZERO: case D_FMIN: case D_FMAX: case D_SIGN: case D_COPYSIGN:
      return 1;
    default: return 60;
  }
}
}  // namespace

long long rematerialise(std::vector<Node>* nodes_io, std::vector<int>* seg_begin_io, const RematOptions& opt, RematStats* stats) {
  const std::vector<Node>& N = *nodes_io;
  const std::vector<int>& sb = *seg_begin_io;
  const int n = static_cast<int>(N.size());
  const int S = static_cast<int>(sb.size()) - 1;
  if (S < 2 || opt.load_cost <= 0) return 0;
  std::vector<int> seg(n, 0);
  for (int s = 0; s < S; ++s)
    for (int k = sb[s]; k < sb[s + 1]; ++k) seg[k] = s;
  // consumers of every value
  std::vector<int> cstart(n + 1, 0), cons;
  {
    std::vector<int> cnt(n + 1, 0);
    auto each = [&](int k, auto&& f) {
      const Node& nd = N[k];
      if (nd.a >= 0) f(nd.a);
      if (nd.b >= 0 && nd.b != nd.a) f(nd.b);
    };
    for (int k = 0; k < n; ++k) each(k, [&](int u) { cnt[u + 1]++; });
    for (int k = 0; k < n; ++k) cstart[k + 1] = cstart[k] + cnt[k + 1];
    cons.resize(cstart[n]);
    std::vector<int> fill(cstart.begin(), cstart.end() - 1);
    for (int k = 0; k < n; ++k) each(k, [&](int u) { cons[fill[u]++] = k; });
  }
  const int W = opt.load_cost, Win = std::max(1, opt.load_cost / 2);
  std::vector<std::vector<int>> clones(S);  // original node ids recomputed inside each segment, increasing
  std::vector<int> vert(n, -1);             // node -> flow vertex of the current network
  std::vector<int> cand, frontier, touched;
  FlowNet net;
  long long total = 0;
  for (int s = 1; s < S; ++s) {
    const int b = sb[s], e = sb[s + 1];
    // candidates: arithmetic ancestors (outside the segment) of its live-ins, nearest first
    cand.clear();
    touched.clear();
    auto visit = [&](int u) {
      if (u < 0 || u >= b || vert[u] != -1 || N[u].kind == K_CONST || N[u].kind == K_OUTPUT) return;
      vert[u] = -2;
      touched.push_back(u);
      cand.push_back(u);
    };
    for (int k = b; k < e; ++k) { visit(N[k].a); visit(N[k].b); }
    const size_t n_direct = cand.size();
    if (n_direct == 0) continue;
    for (size_t q = 0; q < cand.size() && static_cast<int>(cand.size()) < opt.max_candidates; ++q) {
      const int u = cand[q];
      if (N[u].kind != K_ARITH) continue;
      visit(N[u].a);
      visit(N[u].b);
    }
    // instruction-cache budget: the segment with its clones stays within max_segment_weight (at least half of it is
    // always available to clones)
    long long extra_allowed = 0;
    if (opt.max_segment_weight > 0) {
      long long own = 0;
      for (int k = b; k < e; ++k) own += op_weight(N[k]);
      extra_allowed = std::max(opt.max_segment_weight / 2, opt.max_segment_weight - own);
    }
    int Wcur = W, Wincur = Win;
    std::vector<int> chosen;
    for (int attempt = 0; attempt < 3; ++attempt) {
      // network: vertex 0 = source ("evaluated earlier, loaded here when needed"), 1 = sink (this segment)
      net.reset(2);
      for (int u : cand) vert[u] = net.add_node();
      for (int u : cand) {
        const bool arith = N[u].kind == K_ARITH;
        // can this value be recomputed?  only when its operands are themselves candidates, constants or inputs we know
        bool closed = arith;
        if (arith) {
          const int ops[2] = {N[u].a, N[u].b};
          for (int v : ops)
            if (v >= 0 && N[v].kind != K_CONST && vert[v] < 0) closed = false;
        }
        net.add(0, vert[u], closed ? remat_cost(N[u]) : kInf);  // on the sink side = recomputed here
        // loaded (once) when it stays on the source side and something on the sink side reads it
        const int z = net.add_node();
        net.add(vert[u], z, arith ? Wcur : Wincur);
        bool in_seg = false;
        for (int q = cstart[u]; q < cstart[u + 1]; ++q) {
          const int c = cons[q];
          if (c >= b && c < e) in_seg = true;
          else if (c < b && vert[c] >= 0) net.add(z, vert[c], kInf);
        }
        if (in_seg) net.add(z, 1, kInf);
      }
      net.maxflow(0, 1);
      std::vector<char> reach;
      net.reach_to(1, &reach);  // the smallest sink side
      chosen.clear();
      long long extra = 0;
      for (int u : cand)
        if (N[u].kind == K_ARITH && reach[vert[u]]) { chosen.push_back(u); extra += op_weight(N[u]); }
      if (extra_allowed <= 0 || extra <= extra_allowed) break;
      chosen.clear();
      Wcur = std::max(1, Wcur / 2);
      Wincur = std::max(1, Wincur / 2);
    }
    for (int u : touched) vert[u] = -1;
    std::sort(chosen.begin(), chosen.end());
    total += static_cast<long long>(chosen.size());
    clones[s] = chosen;
  }
  if (total == 0) return 0;
  // ---- rebuild: every segment = its clones (in tape order, hence topological) followed by its own nodes
  std::vector<Node> out;
  out.reserve(static_cast<size_t>(n) + total);
  std::vector<int> newid(n, -1), cloneid(n, -1);
  std::vector<int> nsb;
  for (int s = 0; s < S; ++s) {
    nsb.push_back(static_cast<int>(out.size()));
    auto remap = [&](int v) { return v < 0 ? v : (cloneid[v] >= 0 ? cloneid[v] : newid[v]); };
    for (int u : clones[s]) {
      Node nd = N[u];
      nd.a = remap(nd.a);
      nd.b = remap(nd.b);
      cloneid[u] = static_cast<int>(out.size());
      out.push_back(nd);
    }
    for (int k = sb[s]; k < sb[s + 1]; ++k) {
      Node nd = N[k];
      nd.a = remap(nd.a);
      nd.b = remap(nd.b);
      newid[k] = static_cast<int>(out.size());
      out.push_back(nd);
    }
    for (int u : clones[s]) cloneid[u] = -1;
  }
  nsb.push_back(static_cast<int>(out.size()));
  // ---- values nobody reads any more (every reader recomputes them) are dropped
  const int m = static_cast<int>(out.size());
  std::vector<int> uses(m, 0);
  for (int k = 0; k < m; ++k) {
    if (out[k].a >= 0) uses[out[k].a]++;
    if (out[k].b >= 0 && out[k].b != out[k].a) uses[out[k].b]++;
  }
  std::vector<char> dead(m, 0);
  long long dropped = 0;
  for (int k = m - 1; k >= 0; --k) {
    if (out[k].kind != K_ARITH || uses[k] > 0) continue;
    dead[k] = 1;
    ++dropped;
    if (out[k].a >= 0) uses[out[k].a]--;
    if (out[k].b >= 0 && out[k].b != out[k].a) uses[out[k].b]--;
  }
  std::vector<int> fin(m, -1);
  std::vector<Node> packed;
  packed.reserve(m);
  std::vector<int> psb;
  size_t si = 0;
  for (int k = 0; k < m; ++k) {
    while (si < nsb.size() - 1 && nsb[si] == k) { psb.push_back(static_cast<int>(packed.size())); ++si; }
    if (dead[k]) continue;
    Node nd = out[k];
    if (nd.a >= 0) nd.a = fin[nd.a];
    if (nd.b >= 0) nd.b = fin[nd.b];
    fin[k] = static_cast<int>(packed.size());
    packed.push_back(nd);
  }
  while (psb.size() < nsb.size() - 1) psb.push_back(static_cast<int>(packed.size()));
  psb.push_back(static_cast<int>(packed.size()));
  if (stats) { stats->cloned = total; stats->dropped = dropped; }
  nodes_io->swap(packed);
  seg_begin_io->swap(psb);
  return total;
}

}  // namespace ccu
