// Integration test (TEST INFRASTRUCTURE): `f.map(N, "cuda")` through the reference's own public C++ API, in a
// libcasadi.so relinked with the patched Map::create and the new CudaMap (build_integration.py).
// Mirrors the reference's map tests: test/python/function.py:658-757 (test_map_node: all parallelizations +
// AD), :938-1009 (mapaccum), check_serialize (helpers.py:1016-1029) and the error behaviour of map.cpp:49.
//
//   test_cuda_map            full run (needs a CUDA device and libcasadi_cuda.so)
//   test_cuda_map --no-gpu   host-side checks only: dispatch, tape export, loud failure without a device
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <unistd.h>

#include "models.hpp"
#include <cuda_map.hpp>           // CudaMap::lowered_tape (new internal header, casadi_b200/host)
#include <casadi/core/mapsum.hpp>  // MapSum::create (internal header of the reference: the class has no public factory with a parallelization argument)

using namespace casadi;

static int g_fail = 0;
#define CHECK(cond, msg) do { if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, std::string(msg).c_str()); ++g_fail; } } while (0)

static double ulp_dist(double a, double b) {
  if (a == b || (std::isnan(a) && std::isnan(b))) return 0;
  if (!std::isfinite(a) || !std::isfinite(b)) return INFINITY;
  long long x, y;
  std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8);
  if (x < 0) x = static_cast<long long>(0x8000000000000000ull) - x;
  if (y < 0) y = static_cast<long long>(0x8000000000000000ull) - y;
  return std::fabs(static_cast<double>(x - y));
}

struct Buffers {
  std::vector<std::vector<double>> in, out;
  std::vector<const double*> arg;
  std::vector<double*> res;
  std::vector<casadi_int> iw;
  std::vector<double> w;
};

// evaluate F through the buffer API (function.cpp:1708-1738) on given inputs
static int g_expect_flag = 0;
static std::vector<std::vector<double>> eval(const Function& F, const std::vector<std::vector<double>>& in,
                                             const std::vector<bool>& null_in = {}, const std::vector<bool>& null_out = {}) {
  Buffers b;
  b.in = in;
  b.arg.assign(F.sz_arg(), nullptr);
  b.res.assign(F.sz_res(), nullptr);
  b.iw.resize(F.sz_iw());
  b.w.resize(F.sz_w());
  b.out.resize(F.n_out());
  for (casadi_int j = 0; j < F.n_in(); ++j)
    b.arg[j] = (j < (casadi_int)null_in.size() && null_in[j]) ? nullptr : b.in[j].data();
  for (casadi_int j = 0; j < F.n_out(); ++j) {
    b.out[j].assign(F.nnz_out(j), -777.0);
    b.res[j] = (j < (casadi_int)null_out.size() && null_out[j]) ? nullptr : b.out[j].data();
  }
  int flag = F(b.arg.data(), b.res.data(), b.iw.data(), b.w.data(), 0);
  CHECK(flag == g_expect_flag, "evaluation of " + F.name() + " returned " + str(flag));
  return b.out;
}

static std::vector<std::vector<double>> random_inputs(const Function& F, unsigned seed, double lo, double hi) {
  std::mt19937_64 g(seed);
  std::vector<std::vector<double>> in(F.n_in());
  for (casadi_int j = 0; j < F.n_in(); ++j) {
    in[j].resize(F.nnz_in(j));
    for (auto& v : in[j]) v = lo + (hi - lo) * std::generate_canonical<double, 53>(g);
  }
  return in;
}

// max ulp distance over all outputs
static double compare(const std::vector<std::vector<double>>& a, const std::vector<std::vector<double>>& b, double* relerr) {
  double worst = 0;
  *relerr = 0;
  CHECK(a.size() == b.size(), "output count");
  for (size_t j = 0; j < a.size(); ++j) {
    CHECK(a[j].size() == b[j].size(), "output size");
    for (size_t k = 0; k < a[j].size(); ++k) {
      worst = std::max(worst, ulp_dist(a[j][k], b[j][k]));
      *relerr = std::max(*relerr, std::fabs(a[j][k] - b[j][k]) / std::max(1.0, std::fabs(b[j][k])));
    }
  }
  return worst;
}

static void host_side_checks() {
  using namespace ccu_models;
  // the tape export reads the same tape SXFunction::eval runs (sx_function.cpp:72-127)
  Function f = cartpole(4);
  CHECK(f.n_instructions() == 522, "cartpole tape length " + str(f.n_instructions()));
  // "serial" etc. are untouched by the patch; unknown strings still raise (map.cpp:49)
  Function s = f.map(3, "serial");
  CHECK(s.class_name() == "Map", s.class_name());
  bool threw = false;
  try { f.map(3, "no_such_mode"); } catch (std::exception& e) { threw = std::string(e.what()).find("Unknown parallelization") != std::string::npos; }
  CHECK(threw, "unknown parallelization must raise");
  // free variables are rejected when the map is created
  SX x = SX::sym("x"), p = SX::sym("p");
  Function ff("ff", {x}, {x * p}, Dict{{"allow_free", true}});
  threw = false;
  try { ff.map(4, "cuda"); } catch (std::exception& e) { threw = std::string(e.what()).find("free variables") != std::string::npos; }
  CHECK(threw, "free variables must be rejected at creation");
  // an MX function that can neither be expanded (Linsol call) nor lowered node by node
  // (monitor prints: a side effect without a device lowering) fails loudly at creation: no fallback
  {
    Sparsity sp = kkt_sparsity();
    MX K = MX::sym("K", sp), b = MX::sym("b", 60);
    Function g("g", {K, b}, {solve(K, b, "ldl"), b.monitor("b")});
    threw = false;
    try { g.map(4, "cuda"); } catch (std::exception& e) { threw = std::string(e.what()).find("no device lowering") != std::string::npos; }
    CHECK(threw, "MX function with an unsupported node must be rejected");
  }
}

// ---- the MX vocabulary around a Linsol call (widened): a function that holds solve(K, b, "ldl") cannot be
// expanded, so CudaMap lowers it node by node; every node class below must give the bits of the reference's own numeric
// evaluation (f.map(n, "serial")).  The oracle evaluates the lowered tape on the host.
static std::vector<std::vector<double>> eval_tape(const CudaMap::Tape& t, casadi_int n, const std::vector<std::vector<double>>& in);
static void check_bits(const std::vector<std::vector<double>>& got, const std::vector<std::vector<double>>& want, const std::string& what);
static std::vector<std::vector<double>> kkt_like_inputs(const Function& F, casadi_int n, unsigned seed) {
  // input 0: a symmetric banded matrix (diagonal dominant so that LDL is well conditioned), the rest: U(-1, 1)
  std::mt19937_64 g(seed);
  std::vector<std::vector<double>> in(F.n_in());
  for (casadi_int j = 0; j < F.n_in(); ++j) {
    in[j].resize(F.nnz_in(j));
    for (auto& v : in[j]) v = -1 + 2 * std::generate_canonical<double, 53>(g);
  }
  Sparsity sp = Sparsity::banded(6, 1);
  const casadi_int nnz = sp.nnz();
  std::vector<casadi_int> row = sp.get_row(), col = sp.get_col();
  for (casadi_int i = 0; i < n; ++i)
    for (casadi_int k = 0; k < nnz; ++k) {
      double& v = in[0][i * nnz + k];
      if (row[k] == col[k]) v = 4 + v;
      else v = 0.25 * (1 + static_cast<double>((row[k] + col[k] + i) % 3));  // symmetric: depends on row + col only
    }
  return in;
}

static void mx_vocabulary_checks() {
  const casadi_int n = 16;
  Sparsity sp = Sparsity::banded(6, 1);
  MX K = MX::sym("K", sp), b = MX::sym("b", 6), M = MX::sym("M", 3, 6), al = MX::sym("al");
  MX x = solve(K, b, "ldl");
  auto check = [&](const std::string& what, const std::vector<MX>& in, const std::vector<MX>& out, unsigned seed) {
    Function f("voc_" + what, in, out);
    bool expands = true;
    try { f.expand(); } catch (std::exception&) { expands = false; }
    CHECK(!expands, what + ": the case must not be expandable (otherwise it does not test the lowering)");
    Function ref = f.map(n, "serial");
    auto vin = kkt_like_inputs(ref, n, seed);
    try {
      CudaMap::Tape t = CudaMap::lowered_tape(f);
      check_bits(eval_tape(t, n, vin), eval(ref, vin), "MX vocabulary: " + what);
    } catch (std::exception& e) {
      CHECK(false, "MX vocabulary: " + what + " was refused: " + e.what());
    }
    return f;
  };
  // dense products (reference loops), dense and sparse transposes, reductions
  check("dense", {K, b, M}, {mtimes(M, x), mtimes(x.T(), M.T()), mtimes(densify(K), x), mtimes(K.T(), x), dot(x, b), norm_2(x),
                            sumsqr(x), bilin(K, x, b), mmax(x), mmin(K), mmin(x), norm_inf(x), norm_1(x), norm_fro(K),
                            norm_inf(K), norm_1(M)}, 41);
  // logsumexp (numeric evaluation only in the reference; its arg-max is data)
  check("logsumexp", {K, b}, {logsumexp(x), logsumexp(b), logsumexp(x(0)), logsumexp(vertcat(b(0), b(0), x(1)))}, 42);
  // projections, rank-1 update, casts
  check("project", {K, b, al}, {project(K, Sparsity::diag(6)), project(x, Sparsity::dense(6, 1)), rank1(densify(K), al, x, b),
                                sparsity_cast(x, Sparsity::dense(2, 3)), project(K, Sparsity::dense(6, 6)) - 2 * densify(K)}, 43);
  // get / set nonzeros in their vector and slice forms
  {
    MX v = MX::zeros(10, 1);
    v(Slice(2, 8)) = x;
    MX u = MX::zeros(10, 1);
    u(std::vector<casadi_int>{9, 0, 4, 3, 7, 1}) = x;
    MX k2 = K;
    k2.nz(Slice(0, 6)) = x;
    check("nonzeros", {K, b}, {v, u, x(Slice(0, 6, 2)), x.nz(std::vector<casadi_int>{5, 0, 3, 3}), k2, x(Slice(1, 5))}, 44);
  }
  // derivative functions of a function with a linear solve (what Map::get_forward / get_reverse map): the forward and the
  // adjoint sensitivities hold transposed solves, products, projections and add-nonzeros nodes
  {
    Function f("voc_ad", {K, b}, {x(Slice(0, 6, 2)), dot(x, x)});
    for (casadi_int nd : {1, 2}) {
      Function fw = f.forward(nd), rv = f.reverse(nd);
      for (const Function& d : {fw, rv}) {
        Function ref = d.map(n, "serial");
        auto vin = kkt_like_inputs(ref, n, 50 + static_cast<unsigned>(nd));
        try {
          check_bits(eval_tape(CudaMap::lowered_tape(d), n, vin), eval(ref, vin), "MX vocabulary: " + d.name());
        } catch (std::exception& e) {
          CHECK(false, "MX vocabulary: " + d.name() + " was refused: " + e.what());
        }
      }
    }
  }
  // the "tridiag" plugin (casadi/solvers/linsol_tridiag.cpp: the Thomas algorithm): plain, transposed, two right-hand sides,
  // and the derivative functions
  {
    MX B2 = MX::sym("B2", 6, 2);
    Function ft = check("tridiag", {K, b, B2}, {solve(K, b, "tridiag"), solve(K.T(), b, "tridiag"), solve(K, B2, "tridiag"),
                                               mtimes(K, solve(K, b, "tridiag")) - b}, 46);
    for (const Function& d : {ft.forward(1), ft.reverse(1)}) {
      Function ref = d.map(n, "serial");
      auto vin = kkt_like_inputs(ref, n, 47);
      try {
        check_bits(eval_tape(CudaMap::lowered_tape(d), n, vin), eval(ref, vin), "MX vocabulary: " + d.name());
      } catch (std::exception& e) {
        CHECK(false, "MX vocabulary: " + d.name() + " was refused: " + e.what());
      }
    }
    bool threw = false;
    MX Kw = MX::sym("Kw", Sparsity::banded(6, 2));
    Function bad("voc_tri_bad", {Kw, b}, {solve(Kw, b, "tridiag")});
    try { CudaMap::lowered_tape(bad); } catch (std::exception& e) { threw = std::string(e.what()).find("tridiagonal pattern") != std::string::npos; }
    CHECK(threw, "a pattern that is not tridiagonal must be refused by the tridiag lowering");
  }
  // Jacobian and second-order functions of a function with a linear solve (Function::jacobian: multiple right-hand sides,
  // projections; forward-over-reverse as the Hessian-vector products of an NLP with an embedded solve)
  {
    Function f("voc_ad2", {K, b}, {dot(x, x) + bilin(K, x, b), x(Slice(1, 4))});
    for (const Function& d : {f.jacobian(), f.reverse(1).forward(1), f.forward(1).reverse(1)}) {
      Function ref = d.map(n, "serial");
      auto vin = kkt_like_inputs(ref, n, 58);
      try {
        check_bits(eval_tape(CudaMap::lowered_tape(d), n, vin), eval(ref, vin), "MX vocabulary: " + d.name());
      } catch (std::exception& e) {
        CHECK(false, "MX vocabulary: " + d.name() + " was refused: " + e.what());
      }
    }
  }
  // BASELINE config 5 itself: the derivative functions of [x = solve(K, b); r = K*x - b] for both solvers (transposed QR solves)
  for (std::string solver : {"ldl", "qr"}) {
    Function f = ccu_models::kkt_solve(solver);
    Sparsity ksp = ccu_models::kkt_sparsity();
    for (const Function& d : {f.forward(1), f.reverse(1)}) {
      const casadi_int m = 5;
      Function ref = d.map(m, "serial");
      auto vin = random_inputs(ref, 61, -1, 1);
      for (casadi_int i = 0; i < m; ++i) {
        std::vector<double> v = ccu_models::kkt_values(ksp, i);
        std::copy(v.begin(), v.end(), vin[0].begin() + i * ksp.nnz());
      }
      try {
        CudaMap::Tape t = CudaMap::lowered_tape(d);
        auto got = eval_tape(t, m, vin);
        got.resize(ref.n_out());  // (a trailing failure-count output of the QR lowering is not part of the function)
        check_bits(got, eval(ref, vin), "MX vocabulary: " + d.name() + " (" + solver + ")");
      } catch (std::exception& e) {
        CHECK(false, "MX vocabulary: " + d.name() + " (" + solver + ") was refused: " + e.what());
      }
    }
  }
  // maps embedded in the function: g.map(3) and the summed form (MapSum) called on pieces of the solution
  {
    SX p = SX::sym("p", 2), q = SX::sym("q");
    Function g("g", {p, q}, {p * q - p(0) / (p(1) + 3), p(0) + q * p(1)});
    Function G = g.map(3, "serial");
    std::vector<MX> r = G(std::vector<MX>{reshape(x, 2, 3), b(Slice(0, 3)).T()});
    // (Function::map with reductions is Map + HorzRepmat + HorzRepsum, function.cpp:797-818; the MapSum node itself comes
    // from MapSum::create / Function::mapsum)
    Function Gs = MapSum::create("gs", "serial", g, 3, std::vector<bool>{false, true}, std::vector<bool>{true, false});
    bool has_mapsum = Gs.class_name() == "MapSum";  // (wrap_as_needed may put the node inside an MX function)
    for (const std::string& nm : Gs.get_function()) has_mapsum = has_mapsum || Gs.get_function(nm).class_name() == "MapSum";
    CHECK(has_mapsum, "MapSum::create must give a MapSum node, got " + Gs.class_name());
    Function Gr = g.map("gr", "serial", 3, std::vector<casadi_int>{1}, std::vector<casadi_int>{0});
    std::vector<MX> rr = Gr(std::vector<MX>{reshape(x, 2, 3), b(3)});
    std::vector<MX> rs = Gs(std::vector<MX>{reshape(x, 2, 3), b(3)});
    check("maps", {K, b}, {r.at(0), r.at(1), rs.at(0), rs.at(1), rr.at(0), rr.at(1)}, 45);
  }
  // a conditional (Function::conditional -> Switch, switch.cpp:153-211) next to the solve: the reference evaluates the one
  // case the index selects, the lowering all of them, merged by selects on trunc(index) == k; cases with their own patterns
  {
    SX p = SX::sym("p", 6), q = SX::sym("q", 6), ps = SX::sym("ps", Sparsity::diag(3));
    Function f0("f0", {p, q}, {p * q - 1, dot(p, q)});
    Function f1("f1", {p, q}, {p / (q * q + 1), sum1(p) * sum1(q)});
    Function fd("fd", {p, q}, {-p, SX(1, 1)});                     // (a structurally empty second result: projected)
    Function sw = Function::conditional("sw", {f0, f1}, fd);
    MX ind = MX::sym("ind");
    std::vector<MX> r = sw(std::vector<MX>{ind, x, b});
    Function f = check("switch", {K, b, ind}, {r.at(0), r.at(1) + x(0)}, 48);
    // the index values that matter: each case, beyond both ends, and fractions that truncate toward zero
    Function ref = f.map(n, "serial");
    auto vin = kkt_like_inputs(ref, n, 49);
    const double idx[] = {0, 1, 2, -1, 0.7, 1.9, -0.5, -0.0, 1e9, -3.2, 1.0000001, 0.999999};
    for (casadi_int i = 0; i < n; ++i) vin[2][i] = idx[i % 12];
    check_bits(eval_tape(CudaMap::lowered_tape(f), n, vin), eval(ref, vin), "MX vocabulary: switch with chosen index values");
    for (const Function& d : {f.forward(1), f.reverse(1)}) {
      Function dref = d.map(n, "serial");
      auto din = kkt_like_inputs(dref, n, 50);
      for (casadi_int i = 0; i < n; ++i) din[2][i] = idx[i % 12];
      try {
        check_bits(eval_tape(CudaMap::lowered_tape(d), n, din), eval(dref, din), "MX vocabulary: " + d.name());
      } catch (std::exception& e) {
        CHECK(false, "MX vocabulary: " + d.name() + " was refused: " + e.what());
      }
    }
  }
  // lookup tables: interpolant(..., "linear", ...) has no eval_sx, so a function that calls one is lowered; the table
  // entries of a lookup are gathered by selects on the one-hot left index.  1-D and 2-D tables, two values per grid point,
  // the three lookup modes, points inside, on grid points and outside the grid, and the derivative functions
  {
    std::vector<double> g1 = {-1, -0.5, 0, 0.25, 0.75, 1.5, 2}, v1;
    for (size_t i = 0; i < g1.size(); ++i) v1.push_back(std::sin(3 * g1[i]) + 0.1 * static_cast<double>(i));
    std::vector<std::vector<double>> g2 = {{0, 0.5, 1, 2, 4}, {-1, 0, 1, 3}};
    std::vector<double> v2;
    for (size_t j = 0; j < g2[1].size(); ++j)
      for (size_t i = 0; i < g2[0].size(); ++i)
        for (int k = 0; k < 2; ++k) v2.push_back(std::cos(g2[0][i] + 2 * g2[1][j]) * (k + 1) - 0.3 * static_cast<double>(i * j));
    std::vector<double> gu = {0, 0.5, 1, 1.5, 2, 2.5};
    std::vector<double> vu = {1, -2, 0.5, 3, -0.0, 2};
    int cases = 0;
    for (int which = 0; which < 4; ++which) {
      Function L1 = interpolant("lut1", "linear", {g1}, v1, which == 1 ? Dict{{"lookup_mode", std::vector<std::string>{"binary"}}} : Dict());
      Function L2 = interpolant("lut2", "linear", g2, v2, which == 2 ? Dict{{"lookup_mode", std::vector<std::string>{"binary", "linear"}}} : Dict());
      Function Lu = interpolant("lutu", "linear", {gu}, vu, which == 3 ? Dict{{"lookup_mode", std::vector<std::string>{"exact"}}} : Dict());
      MX a = MX::sym("a"), c = MX::sym("c", 2);
      MX y1 = L1(std::vector<MX>{a}).at(0), y2 = L2(std::vector<MX>{c}).at(0), yu = Lu(std::vector<MX>{2 * a + 1}).at(0);
      Function f("lut_case" + str(which), {a, c}, {y1 * y2(0) + yu, y2 + a, sin(y1)});
      // (Function::expand "succeeds" on it, with the interpolant left as a call inside the SX function: not an expansion)
      bool holds_call = false;
      try {
        Function e = f.expand();
        for (casadi_int k = 0; k < e.n_instructions(); ++k) holds_call = holds_call || e.instruction_id(k) == OP_CALL;
      } catch (std::exception&) { holds_call = true; }
      CHECK(holds_call, "a function with a lookup table must not expand to a pure SX tape (otherwise it does not test the lowering)");
      std::vector<Function> fs = {f, f.forward(1), f.reverse(1), f.jacobian()};
      for (const Function& d : fs) {
        const casadi_int nn = 40;
        Function ref = d.map(nn, "serial");
        auto vin = random_inputs(ref, 71 + which, -1.5, 2.5);
        // the first operand on grid points, on the ends and beyond them
        const double special[] = {-1, -0.5, 0, 0.25, 2, -1.25, 2.5, 0.75, 1.5, -0.0};
        for (casadi_int i = 0; i < 10; ++i) vin[0][i] = special[i];
        const double special2[] = {0, -1, 0.5, 0, 4, 3, 5, 4, -1, -2, 2, 1};
        for (casadi_int i = 0; i < 12; ++i) vin[1][20 + i] = special2[i];
        try {
          check_bits(eval_tape(CudaMap::lowered_tape(d), nn, vin), eval(ref, vin), "lookup tables, case " + str(which) + ": " + d.name());
          ++cases;
        } catch (std::exception& e) {
          CHECK(false, "lookup tables, case " + str(which) + ": " + d.name() + " was refused: " + e.what());
        }
      }
    }
    // parametric variants: the table values, the grid, or both are operands -- one table per instance of the map
    for (int which = 0; which < 3; ++which) {
      Function L = which == 0 ? interpolant("lutp0", "linear", {g1}, 1)
                 : which == 1 ? interpolant("lutp1", "linear", std::vector<casadi_int>{7}, v1)
                              : interpolant("lutp2", "linear", std::vector<casadi_int>{5, 4}, 2);
      MX a = MX::sym("a", which == 2 ? 2 : 1), gp = MX::sym("gp", which == 2 ? 9 : 7), cp = MX::sym("cp", which == 2 ? 40 : 7);
      std::vector<MX> largs = {a};
      if (which >= 1) largs.push_back(gp);
      if (which != 1) largs.push_back(cp);
      Function f("lutp_case" + str(which), {a, gp, cp}, {L(largs).at(0) * 2 + gp(0) * cp(0)});
      for (const Function& d : {f, f.forward(1), f.jacobian()}) {
        const casadi_int nn = 30;
        Function ref = d.map(nn, "serial");
        auto vin = random_inputs(ref, 81 + which, -1.5, 2.5);
        for (casadi_int i = 0; i < nn; ++i) {  // a strictly increasing grid per instance
          const casadi_int ngp = which == 2 ? 9 : 7;
          double acc = -1.0 - 0.01 * static_cast<double>(i);
          for (casadi_int j = 0; j < ngp; ++j) {
            if (which == 2 && j == 5) acc = -1.0;
            acc += 0.2 + 0.05 * static_cast<double>((i + 3 * j) % 7);
            vin[1][i * ngp + j] = acc;
          }
        }
        try {
          check_bits(eval_tape(CudaMap::lowered_tape(d), nn, vin), eval(ref, vin), "parametric lookup tables, case " + str(which) + ": " + d.name());
          ++cases;
        } catch (std::exception& e) {
          CHECK(false, "parametric lookup tables, case " + str(which) + ": " + d.name() + " was refused: " + e.what());
        }
      }
    }
    printf("lookup tables (linear interpolants, three lookup modes, parametric values / grids) and their derivative functions: %d functions lowered\n", cases);
  }
  // B-splines: interpolant(..., "bspline", ...) and MX::bspline nodes (constant and parametric coefficients) -- de Boor's
  // recursion over knots gathered by selects; points inside, on knots, on the ends and outside, and the derivative functions
  {
    std::vector<double> g1 = {-1, -0.5, 0, 0.25, 0.75, 1.5, 2}, v1;
    for (size_t i = 0; i < g1.size(); ++i) v1.push_back(std::sin(3 * g1[i]) + 0.1 * static_cast<double>(i));
    std::vector<std::vector<double>> g2 = {{0, 0.5, 1, 2, 4}, {-1, 0, 1, 3, 3.5}};
    std::vector<double> v2;
    for (size_t j = 0; j < g2[1].size(); ++j)
      for (size_t i = 0; i < g2[0].size(); ++i)
        for (int k = 0; k < 2; ++k) v2.push_back(std::cos(g2[0][i] + 2 * g2[1][j]) * (k + 1) - 0.3 * static_cast<double>(i * j));
    int cases = 0;
    for (int which = 0; which < 3; ++which) {
      MX a = MX::sym("a"), c = MX::sym("c", 2), cf = MX::sym("cf", 5);
      std::vector<MX> outs;
      std::vector<MX> ins = {a, c};
      if (which == 0) {
        Function B1 = interpolant("bs1", "bspline", {g1}, v1);
        // (cubic in both directions: the derivative of a degree-1 spline is a degree-0 one, for which the reference itself writes
        // boor[degree-1] out of bounds at a knot)
        Function B2 = interpolant("bs2", "bspline", g2, v2);
        MX y1 = B1(std::vector<MX>{a}).at(0), y2 = B2(std::vector<MX>{c}).at(0);
        outs = {y1 * y2(0), y2 + a, cos(y1)};
      } else if (which == 1) {
        // a node with its own knot vector (a repeated interior knot) and two values per point
        std::vector<std::vector<double>> kn = {{0, 0, 0, 0.3, 0.3, 0.7, 1, 1, 1}};
        std::vector<double> co;
        for (int i = 0; i < 12; ++i) co.push_back(std::sin(0.7 * i) - 0.2 * i);
        outs = {MX::bspline(a, DM(co), kn, std::vector<casadi_int>{2}, 2, Dict()), a * a};
      } else {
        // parametric coefficients
        std::vector<std::vector<double>> kn = {{-1, -1, -1, -1, 0, 1, 1, 1, 1}};
        ins.push_back(cf);
        outs = {MX::bspline(a, cf, kn, std::vector<casadi_int>{3}, 1, Dict()) + c(0)};
        outs.at(0) = outs.at(0) + 0 * cf(0);
      }
      Function f("bs_case" + str(which), ins, outs);
      std::vector<Function> fs = {f, f.forward(1), f.reverse(1), f.jacobian()};
      for (const Function& d : fs) {
        const casadi_int nn = 40;
        Function ref = d.map(nn, "serial");
        auto vin = random_inputs(ref, 91 + which, which == 0 ? -1.2 : -0.2, which == 0 ? 2.2 : 1.2);
        const double sp0[] = {-1, -0.5, 0, 0.25, 2, -1.25, 2.5, 0.75, 1.5, -0.0, 0.3, 0.7, 1, 1.0000001};
        for (casadi_int i = 0; i < 14; ++i) vin[0][i] = sp0[i];
        const double sp1[] = {0, -1, 0.5, 0, 4, 3.5, 5, 4, -1, -2, 2, 1, 4, -1, 0, 3.5};
        for (casadi_int i = 0; i < 16; ++i) vin[1][20 + i] = sp1[i];
        try {
          check_bits(eval_tape(CudaMap::lowered_tape(d), nn, vin), eval(ref, vin), "B-splines, case " + str(which) + ": " + d.name());
          ++cases;
        } catch (std::exception& e) {
          CHECK(false, "B-splines, case " + str(which) + ": " + d.name() + " was refused: " + e.what());
        }
      }
    }
    printf("B-splines (interpolants, nodes with constant and parametric coefficients) and their derivative functions: %d functions lowered\n", cases);
  }
  // an assertion next to the solve: the value passes through and a violated condition makes the evaluation fail (the
  // lowered tape counts the violating instances in its trailing failure output)
  {
    Function f("voc_assert", {K, b}, {x.attachAssert(b(0) < 0.5, "b(0) must stay below 0.5") * 2});
    Function ref = f.map(n, "serial");
    auto vin = kkt_like_inputs(ref, n, 59);
    for (casadi_int i = 0; i < n; ++i) vin[1][i * 6] = -0.25;             // every instance satisfies the condition
    CudaMap::Tape t = CudaMap::lowered_tape(f);
    auto got = eval_tape(t, n, vin);
    CHECK(got.size() == 2 && got[1].size() == static_cast<size_t>(n), "the lowered tape must carry a failure count per instance");
    double bad = 0;
    for (double v : got[1]) bad += v;
    CHECK(bad == 0, "no instance violates the assertion");
    got.resize(1);
    check_bits(got, eval(ref, vin), "MX vocabulary: assertion (satisfied)");
    vin[1][3 * 6] = 0.75;                                                  // instance 3 violates it
    bool threw = false;
    try { eval(ref, vin); } catch (std::exception& e) { threw = std::string(e.what()).find("Assertion error") != std::string::npos; }
    CHECK(threw, "the reference's map raises on a violated assertion");
    got = eval_tape(t, n, vin);
    bad = 0;
    for (double v : got[1]) bad += v;
    CHECK(bad == 1 && got[1][3] == 1, "exactly the violating instance must be counted");
  }
  // still refused, loudly: nodes without a numeric evaluation in the reference, side effects
  {
    bool threw = false;
    Function h("voc_mon", {K, b}, {x.monitor("x")});
    try { CudaMap::lowered_tape(h); } catch (std::exception& e) { threw = std::string(e.what()).find("no device lowering") != std::string::npos; }
    CHECK(threw, "monitor (a printing side effect) must be refused");
  }
  printf("MX vocabulary around a Linsol call: dense / projections / nonzeros / AD / embedded maps lowered\n");
}

// ---- the tape export on the constructs the reference's own function / SX / MX tests build (test/python/function.py,
// sx.py, mx.py): what `f.map(n, "cuda")` hands to the device must evaluate -- here through the oracle, on the host, where
// the transcendentals are the reference's own libm -- to the bits of `f.map(n, "serial")`.
static void export_checks() {
  const casadi_int n = 24;
  int cases = 0;
  auto check = [&](const Function& f, double lo, double hi, unsigned seed) {
    Function ref = f.map(n, "serial");
    auto vin = random_inputs(ref, seed, lo, hi);
    try {
      CudaMap::Tape t = CudaMap::lowered_tape(f);
      check_bits(eval_tape(t, n, vin), eval(ref, vin), "tape export: " + f.name());
      ++cases;
    } catch (std::exception& e) {
      CHECK(false, "tape export: " + f.name() + " was refused: " + e.what());
    }
  };
  SX x = SX::sym("x"), y = SX::sym("y", 2), z = SX::sym("z", 2, 2), v = SX::sym("v", Sparsity::upper(3));
  // test_map_node's function and variations of its signature: sparse and empty operands, repeated / constant / pass-through outputs
  check(Function("sig0", {x, y, z, v}, {mtimes(z, y) + x, sin(y * x).T(), v / x}), 0.5, 1.5, 1);
  check(Function("sig1", {x, SX::sym("e", 0, 0), y}, {x, x, SX(2, 1), y, SX::zeros(Sparsity(3, 3)), 1 + SX::zeros(2, 2)}), -1, 1, 2);
  check(Function("sig2", {v}, {v.T(), mtimes(v, v), SX::triu(mtimes(v.T(), v)), diag(v), v(Slice(), 1)}), -1, 1, 3);
  check(Function("sig3", {z, y}, {SX::solve(z + 3 * SX::eye(2), y), inv(z + 3 * SX::eye(2)), det(z), trace(z), norm_fro(z), sum1(z), sum2(z)}), 0.2, 1, 4);
  // branches and comparisons (sx.py: if_else, logic; calculus.hpp semantics of sign / fmin / fmax / copysign on the edges)
  check(Function("br0", {x, y}, {if_else(x > 0, y(0) * x, y(1) - x), if_else_zero(y(0) <= y(1), x), x < y, x == floor(x), !(x > 0) || (y(0) > 0)}), -1, 1, 5);
  check(Function("br1", {x, y}, {fmin(x, y), fmax(y, x), sign(y), fabs(y) * copysign(x, y(1)), floor(3 * y), ceil(3 * y), fmod(7 * y, x), remainder(7 * y, x)}), -1, 1, 6);
  // powers and roots with constant and variable exponents (sx.py test_pow / constpow simplifications)
  check(Function("pw0", {x, y}, {pow(x, 2), pow(x, 3), pow(x, 0.5), pow(x, -1), pow(x, -2.5), pow(x, y), constpow(x, SX(1.7)), pow(2, y), sqrt(x) * sq(y), 1 / y, x / y / y}), 0.3, 2, 7);
  // the transcendental set
  check(Function("tr0", {x, y}, {exp(y), log(x), sin(y), cos(y), tan(y), asin(y / 3), acos(y / 3), atan(y), atan2(y, x), sinh(y), cosh(y), tanh(y),
                                 asinh(y), acosh(1 + x), atanh(y / 3), erf(y), erfinv(y / 3), log1p(x), expm1(y), hypot(x, y)}), 0.2, 1.5, 8);
  // AD products of a small dynamics function, as Function::forward / reverse / jacobian / hessian build them
  {
    SX s = SX::sym("s", 3), u = SX::sym("u");
    SX rhs = vertcat(s(1), -sin(s(0)) * s(2) - 0.1 * s(1) + u, (u - s(2)) / (1 + s(0) * s(0)));
    Function dyn("dyn", {s, u}, {s + 0.05 * rhs, dot(rhs, rhs)});
    check(dyn, -1, 1, 9);
    check(dyn.forward(2), -1, 1, 10);
    check(dyn.reverse(2), -1, 1, 11);
    check(dyn.jacobian(), -1, 1, 12);
    check(Function("hdyn", {s, u}, {hessian(dot(rhs, rhs), s), gradient(dot(rhs, rhs), s)}), -1, 1, 13);
    // towers: mapaccum, fold, a nested map, an MX wrapper that calls the SX function twice
    check(dyn.mapaccum("acc", 5, std::vector<casadi_int>{0}, std::vector<casadi_int>{0}), -1, 1, 14);
    check(dyn.fold(4), -1, 1, 15);
    check(dyn.map(3, "serial"), -1, 1, 16);
    MX ms = MX::sym("s", 3), mu = MX::sym("u");
    std::vector<MX> r1 = dyn(std::vector<MX>{ms, mu}), r2 = dyn(std::vector<MX>{r1.at(0), mu * r1.at(1)});
    check(Function("wrap", {ms, mu}, {r2.at(0), r1.at(1) - r2.at(1), vertcat(r1.at(0), r2.at(0))(Slice(1, 5))}), -1, 1, 17);
  }
  printf("tape export: %d reference-style functions evaluate to the bits of the serial map\n", cases);
}

// ---- the fixed-step integrator ("rk" plugin, casadi/solvers/runge_kutta.cpp) under a map.
// The oracle (oracle/oracle.c, the checker) evaluates the tape CudaMap lowers the integrator to; the reference's own
// Integrator::eval through f.map(n, "serial") is the expected result, bit for bit.
extern "C" int oracle_map_eval(long long n_instr, const int* op, const int* i0, const int* i1, const int* i2, const double* d,
                               long long n_in, const long long* nnz_in, long long n_out, const long long* nnz_out, long long N,
                               const double* const* arg, double* const* res, double* w);

static std::vector<std::vector<double>> eval_tape(const CudaMap::Tape& t, casadi_int n, const std::vector<std::vector<double>>& in) {
  std::vector<std::vector<double>> out(t.nnz_out.size());
  std::vector<const double*> arg(t.nnz_in.size());
  std::vector<double*> res(t.nnz_out.size());
  std::vector<long long> ni(t.nnz_in.begin(), t.nnz_in.end()), no(t.nnz_out.begin(), t.nnz_out.end());
  for (size_t j = 0; j < arg.size(); ++j) arg[j] = ni[j] ? in.at(j).data() : nullptr;
  for (size_t j = 0; j < res.size(); ++j) { out[j].assign(no[j] * n, -777.0); res[j] = out[j].data(); }
  std::vector<double> w(t.sz_w + 1);
  int flag = oracle_map_eval(static_cast<long long>(t.op.size()), t.op.data(), t.i0.data(), t.i1.data(), t.i2.data(), t.d.data(),
                             static_cast<long long>(ni.size()), ni.data(), static_cast<long long>(no.size()), no.data(), n,
                             arg.data(), res.data(), w.data());
  CHECK(flag == 0, "oracle_map_eval returned " + str(flag));
  return out;
}

// time-dependent ODE with a parameter, a control and a quadrature: x' = ((1-x1^2) x0 - x1 + u + sin t, p x0), q' = x.x + u^2 cos t
// (exact: + - * / only, so that the device result can be compared bit for bit; otherwise the device's sin/cos are within
// 2 ulp of the host's libm and the comparison needs the tolerance of a composite tape)
static Function rk_integrator(const std::string& name, const std::vector<double>& tout, casadi_int nk, bool simplify = false,
                              bool exact = false) {
  SX x = SX::sym("x", 2), p = SX::sym("p"), u = SX::sym("u"), t = SX::sym("t");
  SX ode = vertcat((1 - x(1) * x(1)) * x(0) - x(1) + u + (exact ? t * t / (1 + t) : sin(t)), p * x(0));
  SX quad = dot(x, x) + u * u * (exact ? 1 - t / 3 : cos(t));
  SXDict dae = {{"x", x}, {"p", p}, {"u", u}, {"t", t}, {"ode", ode}, {"quad", quad}};
  Dict opts = {{"number_of_finite_elements", nk}};
  if (simplify) opts["simplify"] = true;
  return integrator(name, "rk", dae, 0.25, tout, opts);
}

static std::vector<std::vector<double>> integrator_inputs(const Function& F, unsigned seed) {
  auto in = random_inputs(F, seed, -0.8, 0.8);
  // signed zeros and a repeated control: q = qf + 1.*q_prev must keep the reference's zero signs
  if (!in[0].empty()) in[0][0] = -0.0;
  return in;
}

static void check_bits(const std::vector<std::vector<double>>& got, const std::vector<std::vector<double>>& want, const std::string& what) {
  CHECK(got.size() >= want.size(), what + ": output count");
  for (size_t j = 0; j < want.size(); ++j) {
    CHECK(got[j].size() == want[j].size(), what + ": size of output " + str(j));
    if (got[j].size() != want[j].size()) continue;
    CHECK(want[j].empty() || std::memcmp(got[j].data(), want[j].data(), want[j].size() * 8) == 0, what + ": output " + str(j) + " differs in bits");
  }
}

static void integrator_lowering_checks() {
  const casadi_int n = 64;
  // one output time, several output times (3 intervals of different length: 7 finite elements become 2 + 3 + 3)
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<double> tout = variant == 0 ? std::vector<double>{1.0} : std::vector<double>{0.4, 0.9, 1.35};
    Function I = rk_integrator("intg" + str(variant), tout, 7);
    CHECK(I.class_name() == "RungeKutta", I.class_name());
    Function ref = I.map(n, "serial");
    auto in = integrator_inputs(ref, 11 + variant);
    auto want = eval(ref, in);
    CudaMap::Tape t = CudaMap::lowered_tape(I);
    check_bits(eval_tape(t, n, in), want, "lowered rk integrator, variant " + str(variant));
    printf("rk integrator variant %d: lowered tape %zu instructions, xf[0] = %.17g, qf[last] = %.17g\n", variant, t.op.size(),
           want[INTEGRATOR_XF][0], want[INTEGRATOR_QF].back());
    // forward sensitivities: the augmented integrator (Integrator::get_forward) inside its MX wrapper
    for (casadi_int nfwd : {1, 3}) {
      Function dI = I.forward(nfwd);
      Function dref = dI.map(n, "serial");
      auto din = integrator_inputs(dref, 23 + variant);
      auto dwant = eval(dref, din);
      CudaMap::Tape dt = CudaMap::lowered_tape(dI);
      check_bits(eval_tape(dt, n, din), dwant, "lowered forward(" + str(nfwd) + ") of the rk integrator, variant " + str(variant));
    }
  }
  {
    Function I = rk_integrator("intg_exact", {0.4, 0.9, 1.35}, 7, false, true);
    Function ref = I.map(n, "serial");
    auto in = integrator_inputs(ref, 17);
    check_bits(eval_tape(CudaMap::lowered_tape(I), n, in), eval(ref, in), "lowered rk integrator, exact-class dynamics");
  }
  // the simplified form (FixedStepIntegrator::create_advanced, integrator.cpp:1894-1950) is a plain MX function that expands
  {
    Function I = rk_integrator("intg_simple", {1.0}, 5, true);
    CHECK(I.class_name() == "MXFunction", I.class_name());
    Function ref = I.map(n, "serial");
    auto in = integrator_inputs(ref, 31);
    check_bits(eval_tape(CudaMap::lowered_tape(I), n, in), eval(ref, in), "simplified rk integrator");
  }
  // a control that repeats with the other zero sign: Integrator::eval keeps the control it holds while next_stop finds no
  // change (-0. == +0.), which shows in the sign of a zero state
  {
    SX x = SX::sym("x"), u = SX::sym("u");
    Function I = integrator("intg_zero", "rk", SXDict{{"x", x}, {"u", u}, {"ode", u}}, 0.0, std::vector<double>{0.5, 1.0, 1.5},
                            Dict{{"number_of_finite_elements", 2}});
    Function ref = I.map(4, "serial");
    auto in = integrator_inputs(ref, 5);
    in[INTEGRATOR_X0] = {-0.0, -0.0, -0.0, 1.0};
    in[INTEGRATOR_U] = {-0.0, 0.0, -0.0,   0.0, -0.0, -0.0,   -0.0, -0.0, 0.0,   -0.0, 0.0, 0.5};
    auto want = eval(ref, in);
    CHECK(std::signbit(want[INTEGRATOR_XF][2]) && !std::signbit(want[INTEGRATOR_XF][3 + 2]),
          "the reference keeps a -0 control across a +0 one (and a +0 one across a -0 one)");
    check_bits(eval_tape(CudaMap::lowered_tape(I), 4, in), want, "rk integrator with controls that repeat with the other zero sign");
  }
  // adjoint sensitivities: the adjoint integrator Integrator::get_reverse creates has backward states; its backward sweep
  // (impulses of the seeds at the output times, retreat through the forward sweep's tape) is replayed with the reference's
  // data-dependent shortcuts -- no impulse for an all-zero seed, no backward integration before the first impulse -- as
  // bit-exact selects.  One and several output times, one and two adjoint directions, instances with all seeds zero, with
  // the seeds of the last output time zero, and with a NaN seed.
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<double> tout = variant == 0 ? std::vector<double>{1.0} : std::vector<double>{0.4, 0.9, 1.35};
    Function I = rk_integrator("intg_adj" + str(variant), tout, 7);
    for (casadi_int nadj : {1, 2}) {
      Function dI = I.reverse(nadj);
      Function ref = dI.map(n, "serial");
      auto in = integrator_inputs(ref, 41 + variant + 10 * static_cast<unsigned>(nadj));
      const casadi_int first_seed = I.n_in() + I.n_out();
      for (casadi_int j = first_seed; j < dI.n_in(); ++j) {
        const casadi_int nz = dI.nnz_in(j);
        for (casadi_int i : {0, 5}) std::fill(in[j].begin() + i * nz, in[j].begin() + (i + 1) * nz, 0.0);  // no seed at all
      }
      {  // instance 7: nothing at the last output time (the backward integration starts one interval earlier)
        const casadi_int j = first_seed + INTEGRATOR_XF, nz = dI.nnz_in(j), nx = 2, nt = static_cast<casadi_int>(tout.size());
        for (casadi_int d = 0; d < nadj; ++d)
          for (casadi_int e = 0; e < nx; ++e) in[j][7 * nz + d * nx * nt + (nt - 1) * nx + e] = 0.0;
        const casadi_int jq = first_seed + INTEGRATOR_QF, nzq = dI.nnz_in(jq);
        for (casadi_int d = 0; d < nadj; ++d) in[jq][7 * nzq + d * nt + (nt - 1)] = 0.0;
        in[j][9 * nz] = std::numeric_limits<double>::quiet_NaN();  // instance 9: a NaN seed is "not zero"
      }
      try {
        CudaMap::Tape t = CudaMap::lowered_tape(dI);
        check_bits(eval_tape(t, n, in), eval(ref, in), "lowered reverse(" + str(nadj) + ") of the rk integrator, variant " + str(variant));
        if (variant == 1 && nadj == 1) printf("rk integrator reverse(1): lowered tape %zu instructions\n", t.op.size());
      } catch (std::exception& e) {
        CHECK(false, "reverse(" + str(nadj) + ") of the rk integrator, variant " + str(variant) + " was refused: " + e.what());
      }
    }
  }
  // second order: forward-over-adjoint (the adjoint integrator augmented with forward sensitivities: stepB also calls
  // fwd<n>_adj<m>_step, :2178-2217) and adjoint-over-forward
  {
    Function I = rk_integrator("intg_foa", {0.5, 1.1}, 5);
    for (int which = 0; which < 3; ++which) {
      // (I.reverse(2).forward(2) cannot be constructed in the reference itself: shape mismatch in Function::call of the augmented integrator)
      Function dI = which == 0 ? I.reverse(1).forward(1) : which == 1 ? I.reverse(1).forward(2) : I.forward(1).reverse(1);
      Function ref = dI.map(n, "serial");
      auto in = integrator_inputs(ref, 61 + which);
      try {
        check_bits(eval_tape(CudaMap::lowered_tape(dI), n, in), eval(ref, in), "lowered second-order derivative " + str(which) + " of the rk integrator");
      } catch (std::exception& e) {
        CHECK(false, "second-order derivative " + str(which) + " of the rk integrator was refused: " + e.what());
      }
    }
  }
}

static void check_close(const std::vector<std::vector<double>>& got, const std::vector<std::vector<double>>& want, double rtol,
                        const std::string& what) {
  double rel = 0;
  compare(std::vector<std::vector<double>>(got.begin(), got.begin() + std::min(got.size(), want.size())), want, &rel);
  CHECK(rel <= rtol, what + ": relative error " + str(rel));
}

// ---- SURVEY 8f-4, second half: the Newton rootfinder (casadi/solvers/newton.cpp) under a map.  Host-side check: the two
// tapes of CudaMap::newton_plan are evaluated by the oracle over plain arrays, driven by the same CudaMap::newton_run
// loop the device uses; expected = the reference's Newton::solve through rf.map(n, "serial"), bit for bit.
struct OracleNewtonBackend : public CudaMap::NewtonBackend {
  const CudaMap::NewtonPlan& P;
  explicit OracleNewtonBackend(const CudaMap::NewtonPlan& p) : P(p) {}
  double* alloc(casadi_int n) override { return new double[n]; }
  void release(double* p) override { delete[] p; }
  void upload(double* dst, const double* src, casadi_int n) override {
    if (src) std::memcpy(dst, src, n * 8); else std::fill(dst, dst + n, 0.0);
  }
  void download(double* dst, const double* src, casadi_int n) override { std::memcpy(dst, src, n * 8); }
  int launch(int which, casadi_int N, const std::vector<const double*>& arg, const std::vector<double*>& res, double* counts) override {
    const CudaMap::Tape& t = P.tape[which];
    std::vector<long long> ni(t.nnz_in.begin(), t.nnz_in.end()), no(t.nnz_out.begin(), t.nnz_out.end());
    std::vector<double> cnt(N * CudaMap::NEWTON_COUNTS), w(t.sz_w + 1);
    std::vector<double*> r(res);
    r.back() = cnt.data();
    int flag = oracle_map_eval(static_cast<long long>(t.op.size()), t.op.data(), t.i0.data(), t.i1.data(), t.i2.data(), t.d.data(),
                               static_cast<long long>(ni.size()), ni.data(), static_cast<long long>(no.size()), no.data(), N,
                               arg.data(), r.data(), w.data());
    for (int k = 0; k < CudaMap::NEWTON_COUNTS; ++k) {
      counts[k] = 0;
      for (casadi_int i = 0; i < N; ++i) counts[k] += cnt[i * CudaMap::NEWTON_COUNTS + k];
    }
    return flag;
  }
};

// the rootfinders of the checks: (0) x^2 - y, the reference's own mapped case (test/python/linearsolver.py:541-549);
// (1) a 2 x 2 system with a parameter and an auxiliary output, exact-class; (2) atan(x) = y/4 from far away (the line
// search must damp the step); (3) as (0) without line search and with too few iterations, failures ignored
static Function newton_case(int which) {
  if (which == 0 || which == 3 || which == 6) {
    MX x = MX::sym("x"), y = MX::sym("y");
    Function f("f", {x, y}, {x * x - y});
    Dict opts = {{"linear_solver", "qr"}};
    if (which == 3) { opts["line_search"] = false; opts["max_iter"] = 3; opts["error_on_fail"] = false; }
    // 6: some instances start at x = 0, where the Jacobian 2x is singular: Newton::solve ignores the failed factorisation
    if (which == 6) { opts["line_search"] = false; opts["max_iter"] = 30; opts["error_on_fail"] = false; }
    return rootfinder("finv" + str(which), "newton", f, opts);
  } else if (which == 1) {
    SX x = SX::sym("x", 2), p = SX::sym("p", 2);
    SX g = vertcat(x(0) * x(0) + x(1) * x(1) - (3 + p(0)), x(0) - x(1) * (1 + p(1) / 4));
    Function f("g2", {x, p}, {g, x(0) * x(1) + p(0)});
    return rootfinder("rf2", "newton", f, Dict{{"max_iter", 60}});
  }
  if (which == 7 || which == 8) {
    // the "fast_newton" plugin on the two-variable system of case 1 (7) and with too few iterations (8)
    SX x = SX::sym("x", 2), p = SX::sym("p", 2);
    SX g = vertcat(x(0) * x(0) + x(1) * x(1) - (3 + p(0)), x(0) - x(1) * (1 + p(1) / 4));
    Function f("g2f", {x, p}, {g, x(0) * x(1) + p(0)});
    Dict opts = {{"max_iter", which == 7 ? 60 : 2}};
    if (which == 8) opts["error_on_fail"] = false;
    return rootfinder("rff" + str(which), "fast_newton", f, opts);
  }
  if (which == 4 || which == 5) {
    // a 4-variable system with a symmetric tridiagonal Jacobian: the gradient of sum_i (x_i^4/4 + x_i^2) + sum_i x_i x_{i+1}/2 - p.x,
    // solved with the "ldl" (4) and the "tridiag" (5) linear solvers
    SX x = SX::sym("x", 4), p = SX::sym("p", 4);
    SX g = x * x * x + 2 * x - p;
    for (int i = 0; i < 3; ++i) { g(i) += x(i + 1) / 2; g(i + 1) += x(i) / 2; }
    Function f("gtri", {x, p}, {g, dot(x, x)});
    return rootfinder("rftri" + str(which), "newton", f, Dict{{"linear_solver", which == 4 ? "ldl" : "tridiag"}, {"max_iter", 60}});
  }
  SX x = SX::sym("x"), y = SX::sym("y");
  Function f("fatan", {x, y}, {atan(x) - y / 4});
  return rootfinder("rfatan", "newton", f, Dict{{"max_iter", 80}});
}

static std::vector<std::vector<double>> newton_inputs(int which, casadi_int n) {
  std::vector<std::vector<double>> in(2);
  std::mt19937_64 g(97 + which);
  auto U = [&](double lo, double hi) { return lo + (hi - lo) * std::generate_canonical<double, 53>(g); };
  if (which == 0 || which == 3 || which == 6) {
    in[0].assign(n, 1.0);
    if (which == 6) for (casadi_int i = 0; i < n; i += 9) in[0][i] = 0.0;
    for (casadi_int i = 0; i < n; ++i) in[1].push_back(n > 1 ? 10.0 * i / (n - 1) : 2.0);  // y = 0: a double root, ~20 iterations
  } else if (which == 1 || which == 7 || which == 8) {
    for (casadi_int i = 0; i < n; ++i) { in[0].push_back(U(0.5, 2.5)); in[0].push_back(U(0.5, 2.5)); in[1].push_back(U(-1, 1)); in[1].push_back(U(-1, 1)); }
  } else if (which == 4 || which == 5) {
    for (casadi_int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) { in[0].push_back(U(-1, 1)); in[1].push_back(U(-3, 3)); }
  } else {
    for (casadi_int i = 0; i < n; ++i) { in[0].push_back(U(-4, 4)); in[1].push_back(U(-3, 3)); }
  }
  return in;
}

static void newton_lowering_checks() {
  const casadi_int n = 200;
  for (int which = 0; which < 9; ++which) {
    Function rf = newton_case(which);
    CHECK(CudaMap::is_newton(rf), rf.class_name());
    Function ref = rf.map(n, "serial");
    auto in = newton_inputs(which, n);
    auto want = eval(ref, in);
    CudaMap::NewtonPlan P = CudaMap::newton_plan(rf);
    OracleNewtonBackend be(P);
    std::vector<std::vector<double>> got(rf.n_out());
    std::vector<const double*> arg(rf.n_in());
    std::vector<double*> res(rf.n_out());
    for (casadi_int j = 0; j < rf.n_in(); ++j) arg[j] = in[j].data();
    for (casadi_int j = 0; j < rf.n_out(); ++j) { got[j].assign(rf.nnz_out(j) * n, -777.0); res[j] = got[j].data(); }
    casadi_int n_failed = -1, n_singular = -1, launches[2] = {0, 0};
    int flag = CudaMap::newton_run(P, n, arg.data(), res.data(), be, &n_failed, &n_singular, launches);
    CHECK(flag == 0, "newton_run flag " + str(flag));
    CHECK(which == 6 ? n_singular > 0 : n_singular == 0, "singular " + str(n_singular));
    // (case 6: the step through the singular factorisation is -inf, the next iterate NaN, and max|F| over NaNs is 0 under
    // std::max: the reference "converges" to NaN, and so does the plan)
    CHECK(which == 3 || which == 8 ? n_failed > 0 : n_failed == 0, "failed instances: " + str(n_failed));
    check_bits(got, want, "Newton rootfinder case " + str(which));
    printf("newton case %d: tapes %zu + %zu instructions, %lld direction + %lld line-search launches, %lld failed, x[last] = %.17g\n", which,
           P.tape[0].op.size(), P.tape[1].op.size(), (long long)launches[0], (long long)launches[1], (long long)n_failed, want[0].back());
  }
}

// The derivative functions of a rootfinder (Rootfinder::get_forward / get_reverse, rootfinder.cpp:318-345, 454-560) do not
// call the solver: they take the nominal solution as an input and are plain MX functions -- the oracle's derivative
// functions, jac_g_x and a (transposed) Linsol solve -- which the lowering handles since its vocabulary is closed under AD.
// So rf.map(n, "cuda").forward(k) / .reverse(k) map them on the device (Map::get_forward); here on the host, bit for bit.
static void newton_derivative_lowering_checks() {
  const casadi_int n = 32;
  for (int which : {1, 2}) {
    Function rf = newton_case(which);
    for (int rev = 0; rev < 2; ++rev) {
      for (casadi_int nd : {1, 3}) {
        Function d = rev ? rf.reverse(nd) : rf.forward(nd);
        Function ref = d.map(n, "serial");
        // nominal inputs as for the solver, the nominal outputs from the solver itself, seeds U(-1, 1)
        auto nom = newton_inputs(which, n);
        auto sol = eval(rf.map(n, "serial"), nom);
        auto vin = random_inputs(ref, 301 + which + 7 * rev, -1, 1);
        for (casadi_int j = 0; j < rf.n_in(); ++j) vin[j] = nom[j];
        for (casadi_int j = 0; j < rf.n_out(); ++j) if (!vin[rf.n_in() + j].empty()) vin[rf.n_in() + j] = sol[j];
        try {
          CudaMap::Tape t = CudaMap::lowered_tape(d);
          auto got = eval_tape(t, n, vin);
          got.resize(ref.n_out());  // (a trailing failure-count output of the QR lowering is not part of the function)
          check_bits(got, eval(ref, vin), "lowered " + d.name() + " of newton case " + str(which));
        } catch (std::exception& e) {
          CHECK(false, "lowering of " + d.name() + " of newton case " + str(which) + " was refused: " + e.what());
        }
      }
    }
  }
  printf("rootfinder derivative functions (forward / reverse, 1 and 3 directions) lowered bit-exactly\n");
}

static void newton_gpu_checks() {
  for (int which : {0, 1, 3}) {  // (case 2 calls atan: device libm, compared with a tolerance below)
    for (casadi_int n : {200, 5000}) {
      Function rf = newton_case(which);
      Function ref = rf.map(n, "serial"), F = rf.map(n, "cuda");
      CHECK(F.class_name() == "CudaMap", F.class_name());
      auto in = newton_inputs(which, n);
      check_bits(eval(F, in), eval(ref, in), "Newton rootfinder case " + str(which) + " under map(" + str(n) + ", cuda)");
    }
  }
  {
    Function rf = newton_case(2);
    Function ref = rf.map(3000, "serial"), F = rf.map(3000, "cuda");
    auto in = newton_inputs(2, 3000);
    check_close(eval(F, in), eval(ref, in), 1e-9, "Newton rootfinder with atan under map(cuda)");
  }
  {  // too few iterations with error_on_fail (the default): the reference raises, so does the device map
    MX x = MX::sym("x"), y = MX::sym("y");
    Function rf = rootfinder("finv_fail", "newton", Function("f", {x, y}, {x * x - y}), Dict{{"max_iter", 2}});
    Function F = rf.map(50, "cuda");
    auto in = newton_inputs(0, 50);
    bool threw = false;
    try { eval(F, in); } catch (std::exception& e) { threw = std::string(e.what()).find("rootfinder process failed") != std::string::npos; }
    CHECK(threw, "a failed rootfinder instance must raise like Rootfinder::eval");
  }
}

static void integrator_gpu_checks() {
  for (casadi_int n : {3, 1000, 70000}) {
    // exact-class dynamics: the device evaluation has the bits of Integrator::eval
    Function I = rk_integrator("intg_gpu", {0.4, 0.9, 1.35}, 7, false, true);
    Function ref = I.map(n, "serial"), F = I.map(n, "cuda");
    CHECK(F.class_name() == "CudaMap", F.class_name());
    auto in = integrator_inputs(ref, 41);
    check_bits(eval(F, in), eval(ref, in), "rk integrator under map(" + str(n) + ", cuda)");
    if (n == 1000) {
      Function dI = I.forward(2);
      Function dref = dI.map(n, "serial"), dF = dI.map(n, "cuda");
      auto din = integrator_inputs(dref, 43);
      check_bits(eval(dF, din), eval(dref, din), "forward(2) of the rk integrator under map(cuda)");
      // and the derivative of the map itself (Map::get_forward -> CudaMap over the augmented integrator)
      Function Fd = F.forward(2), Rd = ref.forward(2);
      auto fin = integrator_inputs(Rd, 47);
      check_bits(eval(Fd, fin), eval(Rd, fin), "forward(2) of map(cuda) of the rk integrator");
      // dynamics with sin/cos of the time: composite tolerance (the device's sin/cos are within 2 ulp of glibc's)
      Function It = rk_integrator("intg_gpu_trig", {0.4, 0.9, 1.35}, 7);
      Function tref = It.map(n, "serial"), tF = It.map(n, "cuda");
      auto tin = integrator_inputs(tref, 53);
      check_close(eval(tF, tin), eval(tref, tin), 1e-11, "rk integrator with sin/cos dynamics under map(cuda)");
    }
  }
  // Device checks of the derivative maps whose lowering is pinned bit for bit on the host (adjoint and second-order
  // sensitivities of the integrator, forward / reverse of the rootfinder) but which no B200 run of this round has seen:
  // opt-in (CCU_TEST_DEVICE_EXTRA=1) until one has.
  if (getenv("CCU_TEST_DEVICE_EXTRA")) {
    const casadi_int n = 1000;
    Function I = rk_integrator("intg_gpu_adj", {0.4, 0.9, 1.35}, 7, false, true);
    for (int which = 0; which < 3; ++which) {
      Function dI = which == 0 ? I.reverse(1) : which == 1 ? I.reverse(2) : I.reverse(1).forward(1);
      Function dref = dI.map(n, "serial"), dF = dI.map(n, "cuda");
      auto din = integrator_inputs(dref, 71 + which);
      check_bits(eval(dF, din), eval(dref, din), "derivative " + str(which) + " (adjoint / second order) of the rk integrator under map(cuda)");
    }
    Function F = I.map(n, "cuda"), ref = I.map(n, "serial");
    Function Fr = F.reverse(1), Rr = ref.reverse(1);
    auto rin = integrator_inputs(Rr, 79);
    check_bits(eval(Fr, rin), eval(Rr, rin), "reverse(1) of map(cuda) of the rk integrator");
    Function rf = newton_case(1);
    for (int rev = 0; rev < 2; ++rev) {
      Function d = rev ? rf.reverse(1) : rf.forward(1);
      Function dref = d.map(n, "serial"), dF = d.map(n, "cuda");
      auto nom = newton_inputs(1, n);
      auto sol = eval(rf.map(n, "serial"), nom);
      auto vin = random_inputs(dref, 83 + rev, -1, 1);
      for (casadi_int j = 0; j < rf.n_in(); ++j) vin[j] = nom[j];
      for (casadi_int j = 0; j < rf.n_out(); ++j) if (!vin[rf.n_in() + j].empty()) vin[rf.n_in() + j] = sol[j];
      check_bits(eval(dF, vin), eval(dref, vin), std::string(rev ? "reverse" : "forward") + "(1) of the Newton rootfinder under map(cuda)");
    }
    printf("extra device checks done\n");
  }
}

static void no_gpu_checks() {
  using namespace ccu_models;
  bool threw = false;
  std::string msg;
  try { cartpole(1).map(4, "cuda"); } catch (std::exception& e) { threw = true; msg = e.what(); }
  CHECK(threw, "without a CUDA device the map must fail at creation");
  CHECK(msg.find("no CPU fallback") != std::string::npos || msg.find("Cannot load") != std::string::npos, msg);
  printf("no-gpu message: %s\n", msg.substr(0, 300).c_str());
}

static void gpu_checks() {
  using namespace ccu_models;
  // ---- test_map_node (function.py:658-696): all parallelizations, n=2 and a larger n
  Function f = map_node_fun();
  for (casadi_int n : {2, 50, 1000}) {
    Function ref = f.map(n, "serial");
    auto in = random_inputs(ref, 7, 0.1, 1.0);
    auto want = eval(ref, in);
    for (std::string par : {"openmp", "thread", "cuda"}) {
      if (par == "thread" && n > 50) continue;
      Function F = f.map(n, par);
      if (par == "cuda") {
        CHECK(F.class_name() == "CudaMap", F.class_name());
        CHECK(F.is_a("Map", true), "CudaMap is_a Map");
        CHECK(F.name() == "cudamap" + str(n) + "_f", F.name());
        for (casadi_int j = 0; j < F.n_in(); ++j) CHECK(F.sparsity_in(j) == ref.sparsity_in(j), "sparsity_in");
        for (casadi_int j = 0; j < F.n_out(); ++j) CHECK(F.sparsity_out(j) == ref.sparsity_out(j), "sparsity_out");
      }
      auto got = eval(F, in);
      double rel;
      double u = compare(got, want, &rel);
      // outputs 0 and 2 are exact-class (mtimes+add, division); output 1 is sin(): <= 2 ulp
      double r0, r2;
      std::vector<std::vector<double>> g0{got[0], got[2]}, w0{want[0], want[2]};
      CHECK(compare(g0, w0, &r0) == 0, par + ": exact-class outputs must be bit-identical");
      CHECK(u <= 2, par + ": sin output differs by " + str(u) + " ulp");
      (void)r2;
    }
  }
  // ---- null argument / null result (sx_function.cpp:116-117) through the map
  {
    casadi_int n = 20;
    Function ref = f.map(n, "serial"), F = f.map(n, "cuda");
    auto in = random_inputs(ref, 8, 0.1, 1.0);
    auto want = eval(ref, in, {false, true, false, false}, {false, true, false});
    auto got = eval(F, in, {false, true, false, false}, {false, true, false});
    double rel;
    CHECK(compare(got, want, &rel) == 0, "null arg/res");
    CHECK(got[1][0] == -777.0, "null result must not be written");
  }
  // ---- derivatives of the map stay on the device: Map::get_forward/get_reverse map the derivative function
  //      with parallelization() (map.cpp:226,280)
  {
    casadi_int n = 6;
    Function ref = f.map(n, "serial"), F = f.map(n, "cuda");
    for (int rev = 0; rev < 2; ++rev) {
      Function dref = rev ? ref.reverse(2) : ref.forward(2);
      Function dF = rev ? F.reverse(2) : F.forward(2);
      std::vector<std::string> calls;
      bool has_cuda = false;
      for (auto& nm : dF.get_function()) has_cuda = has_cuda || dF.get_function(nm).class_name() == "CudaMap";
      CHECK(has_cuda, std::string(rev ? "reverse" : "forward") + " of a cuda map must call a CudaMap");
      // ... and nothing else: the direction-major <-> instance-major permutations (map.cpp:231-264, 285-318) are folded
      // into the device map's chunk copies, so the derivative function is Input (+ empty constants) -> Call -> Output only
      casadi_int n_perm = 0;
      for (casadi_int k = 0; k < dF.n_instructions(); ++k) {
        casadi_int o = dF.instruction_id(k);
        n_perm += (o != OP_INPUT && o != OP_OUTPUT && o != OP_CALL && o != OP_CONST);  // (constants: the empty nominal outputs)
      }
      CHECK(n_perm == 0, std::string(rev ? "reverse" : "forward") + "(2) of a cuda map still has " + str(n_perm) + " glue nodes");
      Function dG = Function::deserialize(dF.serialize());
      auto in0 = random_inputs(dref, 19 + rev, 0.1, 1.0);
      double rel0;
      CHECK(compare(eval(dG, in0), eval(dF, in0), &rel0) == 0, "deserialized derivative map differs");
      auto in = random_inputs(dref, 9 + rev, 0.1, 1.0);
      auto want = eval(dref, in), got = eval(dF, in);
      double rel;
      compare(got, want, &rel);
      CHECK(rel <= 1e-13, std::string(rev ? "reverse" : "forward") + " sensitivities rel err " + str(rel));
    }
  }
  // ---- serialization round trip (check_serialize, helpers.py:1016-1029; Map::deserialize map.cpp:110-122)
  {
    Function F = f.map(5, "cuda");
    Function G = Function::deserialize(F.serialize());
    CHECK(G.class_name() == "CudaMap", "deserialized class " + G.class_name());
    auto in = random_inputs(F, 11, 0.1, 1.0);
    double rel;
    CHECK(compare(eval(G, in), eval(F, in), &rel) == 0, "deserialized map differs");
  }
  // ---- BASELINE config 0: cart-pole RK4, N = 1e5, against the serial map
  {
    casadi_int n = 100000;
    Function c = cartpole(4);
    Function ref = c.map(n, "serial"), F = c.map(n, "cuda");
    auto in = random_inputs(ref, 1, -0.5, 0.5);
    double rel;
    compare(eval(F, in), eval(ref, in), &rel);
    CHECK(rel <= 1e-12, "cartpole N=1e5 rel err " + str(rel));
  }
  // ---- exact-class tape (rocket hess_lag): bit-identical to the serial map
  {
    casadi_int n = 64;
    Function h = rocket_hess_lag(20);
    Function ref = h.map(n, "serial"), F = h.map(n, "cuda");
    auto in = random_inputs(ref, 3, 0.5, 1.5);
    double rel;
    CHECK(compare(eval(F, in), eval(ref, in), &rel) == 0, "hess_lag must be bit-identical");
  }
  // ---- thread-capped overload Function::map(n, par, max_num_threads) (function.cpp:829-858): it nests
  //      f.map(d, "serial").map(T, "cuda"); CudaMap flattens the inner map to d*T device instances
  {
    Function c = cartpole(4);
    for (casadi_int n : {96, 100}) {   // divisible by 8 / not divisible (helper function around the base map)
      Function ref = c.map(n, "serial"), F = c.map(n, "cuda", 8);
      auto in = random_inputs(ref, 21, -0.5, 0.5);
      double rel;
      compare(eval(F, in), eval(ref, in), &rel);
      CHECK(rel <= 1e-12, "map(n,'cuda',8) n=" + str(n) + " rel err " + str(rel));
    }
    Function nested = c.map(12, "serial").map(8, "cuda");
    CHECK(nested.class_name() == "CudaMap", nested.class_name());
  }
  // ---- MapSum with "cuda" (mapsum.cpp:31-54 only knows "serial"): reduced inputs are one instance, reduced outputs
  //      are summed on the device; derivatives stay on the device (mapsum.cpp:304,367 use parallelization())
  {
    casadi_int n = 3000;
    Function g = mc_leaf();
    std::vector<bool> rin{false, true}, rout{false, true};
    Function ref = MapSum::create("ms_ref", "serial", g, n, rin, rout);
    Function F = MapSum::create("ms_cuda", "cuda", g, n, rin, rout);
    {  // MapSum::create returns an MX wrapper around the node (wrap_as_needed, mapsum.cpp:50), for "serial" as well
      bool found = false;
      for (auto& nm : F.get_function()) found = found || F.get_function(nm).class_name() == "CudaMapSum";
      CHECK(found, "MapSum::create(..., \"cuda\") must embed a CudaMapSum");
    }
    for (casadi_int j = 0; j < F.n_in(); ++j) CHECK(F.sparsity_in(j) == ref.sparsity_in(j), "mapsum sparsity_in");
    for (casadi_int j = 0; j < F.n_out(); ++j) CHECK(F.sparsity_out(j) == ref.sparsity_out(j), "mapsum sparsity_out");
    auto in = random_inputs(ref, 31, -1, 1);
    double rel;
    auto want = eval(ref, in), got = eval(F, in);
    std::vector<std::vector<double>> g0{got[0]}, w0{want[0]};
    compare(g0, w0, &rel);
    CHECK(rel <= 1e-13, "mapsum: mapped output rel err " + str(rel));    // sin() in the leaf
    compare(got, want, &rel);
    CHECK(rel <= 1e-12, "mapsum: summed output rel err " + str(rel));    // tree sum vs sequential sum of 3000 terms
    Function G = Function::deserialize(F.serialize());
    {
      bool found = false;
      for (auto& nm : G.get_function()) found = found || G.get_function(nm).class_name() == "CudaMapSum";
      CHECK(found, "deserialized function must embed a CudaMapSum");
    }
    CHECK(compare(eval(G, in), got, &rel) == 0, "deserialized mapsum differs");
    Function dref = ref.forward(1), dF = F.forward(1);
    auto din = random_inputs(dref, 32, -1, 1);
    compare(eval(dF, din), eval(dref, din), &rel);
    CHECK(rel <= 1e-12, "mapsum forward rel err " + str(rel));
  }
  // ---- MapSum("cuda") over a function that is itself a Map: a reduced input is one WHOLE instance of f_ (d inner
  //      blocks) and a reduced output is the sum of whole f_ outputs (mapsum.cpp:154-186) -- the inner map must not
  //      be flattened into n*d device instances here
  {
    Function g = mc_leaf().map(3, "serial");  // x[4x3], w[2x3] -> xn[4x3], cost[1x3]
    casadi_int n = 50;
    std::vector<bool> rin{false, true}, rout{false, true};
    Function ref = MapSum::create("msn_ref", "serial", g, n, rin, rout);
    Function F = MapSum::create("msn_cuda", "cuda", g, n, rin, rout);
    for (casadi_int j = 0; j < F.n_out(); ++j) CHECK(F.sparsity_out(j) == ref.sparsity_out(j), "nested mapsum sparsity_out");
    auto in = random_inputs(ref, 41, -1, 1);
    double rel;
    auto want = eval(ref, in), got = eval(F, in);
    CHECK(got[1].size() == 3, "nested mapsum: reduced output holds one whole f_ output");
    compare(got, want, &rel);
    CHECK(rel <= 1e-12, "mapsum over a nested map rel err " + str(rel));
  }
  // ---- second order and Jacobians of the mapped function through the reference API, as checkfunction does
  //      (test/python/helpers.py:520-688: jacobian, forward-over-reverse, hessian of a scalarised output)
  {
    casadi_int n = 4;
    Function ref = f.map(n, "serial"), F = f.map(n, "cuda");
    for (int which = 0; which < 3; ++which) {
      Function dref, dF;
      std::string what;
      if (which == 0) { dref = ref.jacobian(); dF = F.jacobian(); what = "jacobian"; }
      if (which == 1) { dref = ref.reverse(1).forward(2); dF = F.reverse(1).forward(2); what = "forward-over-reverse"; }
      if (which == 2) { dref = ref.forward(1).reverse(1); dF = F.forward(1).reverse(1); what = "reverse-over-forward"; }
      for (casadi_int j = 0; j < dF.n_out(); ++j) CHECK(dF.sparsity_out(j) == dref.sparsity_out(j), what + " sparsity_out");
      auto in = random_inputs(dref, 50 + which, 0.1, 1.0);
      double rel;
      compare(eval(dF, in), eval(dref, in), &rel);
      CHECK(rel <= 1e-12, what + " of the cuda map rel err " + str(rel));
    }
    // Hessian of a scalar function of the mapped outputs (MX wrapper calling the map)
    for (std::string par : {"cuda"}) {
      std::vector<MX> a;
      for (casadi_int j = 0; j < ref.n_in(); ++j) a.push_back(MX::sym("a" + str(j), ref.sparsity_in(j)));
      auto scal = [&](const Function& M) {
        std::vector<MX> r = M(a);
        MX s = 0;
        for (auto& e : r) s += sumsqr(e);
        MX xv = vec(a[1]);
        Function S("S", a, {s});
        return S.factory("H", S.name_in(), {"hess:" + S.name_out(0) + ":" + S.name_in(1) + ":" + S.name_in(1)});
      };
      Function Href = scal(ref), Hc = scal(F);
      auto in = random_inputs(Href, 60, 0.1, 1.0);
      double rel;
      compare(eval(Hc, in), eval(Href, in), &rel);
      CHECK(rel <= 1e-11, "hessian through the cuda map rel err " + str(rel));
    }
  }
  // ---- mapaccum tower (function.py:938-1009): an MXFunction that CudaMap expands to one SX tape
  {
    Function acc = mc_leaf().mapaccum(10);
    CHECK(acc.class_name() == "MXFunction", acc.class_name());
    casadi_int n = 300;
    Function ref = acc.map(n, "serial"), F = acc.map(n, "cuda");
    auto in = random_inputs(ref, 4, -1, 1);
    double rel;
    compare(eval(F, in), eval(ref, in), &rel);
    CHECK(rel <= 1e-13, "mapaccum rel err " + str(rel));
  }
  // ---- map with reductions (function.cpp:797-818): HorzRepmat/HorzRepsum around the Map node
  {
    casadi_int n = 40;
    Function g = mc_leaf();
    Function ref = g.map("r", "serial", n, std::vector<casadi_int>{0}, std::vector<casadi_int>{1});
    Function F = g.map("r", "cuda", n, std::vector<casadi_int>{0}, std::vector<casadi_int>{1});
    auto in = random_inputs(ref, 5, -1, 1);
    double rel;
    compare(eval(F, in), eval(ref, in), &rel);
    CHECK(rel <= 1e-13, "map with reduce_in/out rel err " + str(rel));
  }
}

// ---- BASELINE config 5: MX function [x = solve(K,b,solver); r = K*x-b] -- not expandable; CudaMap
//      lowers it node by node, the Linsol call by tracing casadi_ldl / casadi_qr over the shared pattern
static void kkt_checks() {
  using namespace ccu_models;
  Sparsity sp = kkt_sparsity();
  for (std::string solver : {"ldl", "qr"}) {
    Function f = kkt_solve(solver);
    casadi_int n = 500;
    Function ref = f.map(n, "serial"), F = f.map(n, "cuda");
    CHECK(F.class_name() == "CudaMap", F.class_name());
    std::vector<std::vector<double>> in(2);
    std::mt19937_64 g(12);
    for (casadi_int i = 0; i < n; ++i) {
      std::vector<double> v = kkt_values(sp, i);
      in[0].insert(in[0].end(), v.begin(), v.end());
      for (int k = 0; k < 60; ++k) in[1].push_back(-1 + 2 * std::generate_canonical<double, 53>(g));
    }
    double rel;
    CHECK(compare(eval(F, in), eval(ref, in), &rel) == 0, "kkt_" + solver + ": x and r must be bit-identical to the serial map");
    // the derivative maps of config 5 stay on the device: Map::get_forward / get_reverse map f.forward(1) / f.reverse(1),
    // MX functions with (transposed) solves, products, projections and add-nonzeros nodes, lowered node by node
    for (int rev = 0; rev < 2; ++rev) {
      // (the specialised kernels of these 255-register tapes take a minute of NVRTC per map: the interpreter runs the cases
      // by default, CCU_TEST_KKT_AD_JIT=1 specialises the ldl / forward one -- green on a B200 either way, g20)
      const bool interp = !(solver == "ldl" && rev == 0 && getenv("CCU_TEST_KKT_AD_JIT") != nullptr);
      if (interp) setenv("CCU_MODE", "interp", 1);
      Function dref = rev ? ref.reverse(1) : ref.forward(1), dF = rev ? F.reverse(1) : F.forward(1);
      if (interp) unsetenv("CCU_MODE");
      bool has_cuda = false;
      for (const std::string& nm : dF.get_function()) has_cuda = has_cuda || dF.get_function(nm).is_a("CudaMap", true);
      CHECK(has_cuda || dF.is_a("CudaMap", true), std::string(rev ? "reverse" : "forward") + " of the kkt cuda map must call a CudaMap");
      auto din = random_inputs(dref, 71 + rev, -1, 1);
      std::copy(in[0].begin(), in[0].end(), din[0].begin());
      CHECK(compare(eval(dF, din), eval(dref, din), &rel) == 0,
            "kkt_" + solver + std::string(rev ? " reverse(1)" : " forward(1)") + " of the cuda map must be bit-identical to the serial map's");
    }
    if (solver == "qr") {
      // a singular instance makes LinsolQr::nfact fail (linsol_qr.cpp:146-163) and the map return 1
      for (casadi_int k = 0; k < sp.nnz(); ++k) in[0][7 * sp.nnz() + k] = 0;
      auto fails = [&](const Function& fn) {
        int before = g_fail;
        bool failed = false;
        g_expect_flag = 1;
        try { eval(fn, in); failed = g_fail == before; } catch (std::exception&) { failed = true; }
        g_expect_flag = 0;
        g_fail = before;
        return failed;
      };
      CHECK(fails(ref), "the serial map must fail on a singular system");
      CHECK(fails(F), "the cuda map must fail on a singular system like the serial map");
    }
  }
}

int main(int argc, char** argv) {
  bool no_gpu = argc > 1 && std::string(argv[1]) == "--no-gpu";
  {  // Linsol plugins live next to libcasadi.so: <exe dir>/../lib
    char buf[4096];
    ssize_t len = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    if (len > 0) {
      std::string p(buf, len);
      p = p.substr(0, p.rfind('/'));
      GlobalOptions::setCasadiPath(p.substr(0, p.rfind('/')) + "/lib");
    }
  }
  try {
    host_side_checks();
    integrator_lowering_checks();
    newton_lowering_checks();
    newton_derivative_lowering_checks();
    mx_vocabulary_checks();
    export_checks();
    if (no_gpu) no_gpu_checks(); else { gpu_checks(); kkt_checks(); integrator_gpu_checks(); newton_gpu_checks(); }
  } catch (std::exception& e) {
    printf("FAIL: unexpected exception: %s\n", e.what());
    return 1;
  }
  printf(g_fail ? "%d check(s) FAILED\n" : "integration ok%.0d\n", g_fail);
  return g_fail ? 1 : 0;
}
