// Fixture generator (TEST INFRASTRUCTURE).  Links against the unmodified reference built by
// oracle/build_ref.py, builds the BASELINE config models (oracle/models.hpp), and dumps
//   <name>.tape : the SXFunction instruction tape read through the public accessors
//                 Function::n_instructions/instruction_id/_input/_output/_constant
//                 (/root/reference/casadi/core/function.hpp:1114-1138, sx_function.hpp:187-231)
//   <name>.case : seeded AoS inputs + the outputs of f.map(N,"serial") (map.cpp:141-157), the parity oracle
// Raw files are converted to compressed .npz under tests/golden/ by oracle/make_golden.py.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <limits>
#include <random>
#include <functional>
#include <unistd.h>
#include "models.hpp"

using namespace casadi;

static std::string g_outdir;

static void wr(FILE* f, const void* p, size_t n) {
  if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(1); }
}
static void wr64(FILE* f, long long v) { wr(f, &v, 8); }

// opcodes from calculus.hpp:60-218 that carry non-slot operands
enum { K_OP_CONST = 44, K_OP_INPUT = 45, K_OP_OUTPUT = 46, K_OP_PARAMETER = 47, K_OP_CALL = 48 };

static void dump_tape(const Function& f, const std::string& name) {
  casadi_assert(f.is_a("SXFunction"), name + ": not an SXFunction");
  long long n = f.n_instructions();
  std::vector<int> op(n), i0(n, 0), i1(n, 0), i2(n, 0);
  std::vector<double> d(n, 0.0);
  for (long long k = 0; k < n; ++k) {
    op[k] = static_cast<int>(f.instruction_id(k));
    std::vector<casadi_int> in = f.instruction_input(k), out = f.instruction_output(k);
    if (op[k] == K_OP_CONST) {
      i0[k] = out.at(0); d[k] = f.instruction_constant(k);
    } else if (op[k] == K_OP_INPUT) {
      i0[k] = out.at(0); i1[k] = in.at(0); i2[k] = in.at(1);
    } else if (op[k] == K_OP_OUTPUT) {
      i0[k] = out.at(0); i2[k] = out.at(1); i1[k] = in.at(0);
    } else if (op[k] == K_OP_CALL || op[k] == K_OP_PARAMETER) {
      casadi_error(name + ": tape has OP_CALL/OP_PARAMETER");
    } else {
      i0[k] = out.at(0); i1[k] = in.at(0); i2[k] = in.size() > 1 ? in.at(1) : in.at(0);
    }
  }
  FILE* fp = fopen((g_outdir + "/" + name + ".tape").c_str(), "wb");
  wr(fp, "CCUTAPE1", 8);
  wr64(fp, n); wr64(fp, f.sz_w()); wr64(fp, f.n_in()); wr64(fp, f.n_out());
  for (casadi_int j = 0; j < f.n_in(); ++j) wr64(fp, f.nnz_in(j));
  for (casadi_int j = 0; j < f.n_out(); ++j) wr64(fp, f.nnz_out(j));
  wr(fp, op.data(), 4 * n); wr(fp, i0.data(), 4 * n); wr(fp, i1.data(), 4 * n); wr(fp, i2.data(), 4 * n);
  wr(fp, d.data(), 8 * n);
  fclose(fp);
  printf("%-18s n_instr=%lld sz_w=%lld in=", name.c_str(), n, (long long)f.sz_w());
  for (casadi_int j = 0; j < f.n_in(); ++j) printf("%lld ", (long long)f.nnz_in(j));
  printf("out=");
  for (casadi_int j = 0; j < f.n_out(); ++j) printf("%lld ", (long long)f.nnz_out(j));
  printf("\n");
}

typedef std::function<double(int j, long long inst, long long k)> Filler;

// evaluate F = f.map(N, "serial") through the buffer API and write inputs + outputs
static void dump_case(const Function& f, const std::string& name, long long N, const Filler& fill,
                      const std::string& casename = "") {
  Function F = f.map(N, "serial");
  std::vector<std::vector<double>> in(f.n_in()), out(f.n_out());
  for (casadi_int j = 0; j < f.n_in(); ++j) {
    in[j].resize(N * f.nnz_in(j));
    for (long long i = 0; i < N; ++i)
      for (long long k = 0; k < f.nnz_in(j); ++k) in[j][i * f.nnz_in(j) + k] = fill(j, i, k);
  }
  std::vector<const double*> arg(F.sz_arg(), nullptr);
  std::vector<double*> res(F.sz_res(), nullptr);
  std::vector<casadi_int> iw(F.sz_iw());
  std::vector<double> w(F.sz_w());
  for (casadi_int j = 0; j < f.n_in(); ++j) arg[j] = in[j].data();
  for (casadi_int j = 0; j < f.n_out(); ++j) { out[j].assign(N * f.nnz_out(j), -777.0); res[j] = out[j].data(); }
  int flag = F(arg.data(), res.data(), iw.data(), w.data(), 0);
  casadi_assert(flag == 0, "eval failed");
  std::string cn = casename.empty() ? name : casename;
  FILE* fp = fopen((g_outdir + "/" + cn + ".case").c_str(), "wb");
  wr(fp, "CCUCASE1", 8);
  wr64(fp, N); wr64(fp, f.n_in()); wr64(fp, f.n_out());
  for (casadi_int j = 0; j < f.n_in(); ++j) wr64(fp, f.nnz_in(j));
  for (casadi_int j = 0; j < f.n_out(); ++j) wr64(fp, f.nnz_out(j));
  for (auto& v : in) wr(fp, v.data(), 8 * v.size());
  for (auto& v : out) wr(fp, v.data(), 8 * v.size());
  fclose(fp);
}

struct Rng {
  std::mt19937_64 g;
  explicit Rng(unsigned long long s) : g(s) {}
  double u(double a, double b) { return a + (b - a) * std::generate_canonical<double, 53>(g); }
  double n() { std::normal_distribution<double> d(0, 1); return d(g); }
};

int main(int argc, char** argv) {
  g_outdir = argc > 1 ? argv[1] : ".";
  {  // Linsol plugins live next to libcasadi.so: <exe dir>/../lib
    char buf[4096];
    ssize_t len = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    if (len > 0) {
      std::string p(buf, len);
      p = p.substr(0, p.rfind('/'));
      GlobalOptions::setCasadiPath(p.substr(0, p.rfind('/')) + "/lib");
    }
  }
  using namespace ccu_models;
  const double nan = std::numeric_limits<double>::quiet_NaN(), inf = std::numeric_limits<double>::infinity();

  // config 0 --------------------------------------------------------------------------------------
  {
    Function f = cartpole(4);
    dump_tape(f, "cartpole");
    Rng r(1);
    dump_case(f, "cartpole", 1000, [&](int j, long long, long long) { return j == 0 ? r.u(-0.5, 0.5) : r.u(-1, 1); });
    Function f1 = cartpole(1);
    dump_tape(f1, "cartpole1");
    Rng r1(11);
    dump_case(f1, "cartpole1", 333, [&](int j, long long, long long) { return j == 0 ? r1.u(-0.5, 0.5) : r1.u(-1, 1); });
  }
  // config 1 --------------------------------------------------------------------------------------
  {
    const double hover = 1.2 * 9.81 / 4;
    Function F = quadrotor(20);
    Rng r(2);
    auto fill = [&](int j, long long, long long) {
      if (j == 0) return r.u(-0.3, 0.3);
      if (j == 1) return hover * (1 + r.u(-0.1, 0.1));
      return r.u(-1, 1);  // seeds (and nominal-output dummy inputs) of fwd/adj functions
    };
    dump_tape(F, "quad");
    dump_case(F, "quad", 257, fill);
    Function Ff = F.forward(1);
    dump_tape(Ff, "quad_fwd");
    dump_case(Ff, "quad_fwd", 129, fill);
    Function Fr = F.reverse(1);
    dump_tape(Fr, "quad_adj");
    dump_case(Fr, "quad_adj", 129, fill);
    Function J = F.jacobian();
    dump_tape(J, "quad_jac");
    dump_case(J, "quad_jac", 97, fill);
    Function F1 = quadrotor(1);
    dump_tape(F1, "quad1");
    dump_case(F1, "quad1", 257, fill);
    Function J1 = F1.jacobian();
    dump_tape(J1, "quad1_jac");
    dump_case(J1, "quad1_jac", 129, fill);
  }
  // config 2 --------------------------------------------------------------------------------------
  {
    const int K = 20;
    Function H = rocket_hess_lag(K);
    dump_tape(H, "rocket_hess");
    // nominal trajectory: straight-line descent from r0 to origin, hover-ish thrust
    std::vector<double> xnom(1 + 7 + K * 10);
    const double tf = 8.0, g = 3.71, m0 = 1000.0, alpha = 5e-4;
    const double r0[3] = {200, 100, 1500}, v0[3] = {-10, 5, -80};
    xnom[0] = tf;
    auto state = [&](int k, double* s) {
      double t = static_cast<double>(k) / K;
      for (int a = 0; a < 3; ++a) { s[a] = r0[a] * (1 - t) * (1 - t); s[3 + a] = v0[a] * (1 - t); }
      s[6] = m0 - 30.0 * t;
    };
    state(0, &xnom[1]);
    for (int k = 0; k < K; ++k) {
      xnom[8 + k * 10 + 0] = 50.0; xnom[8 + k * 10 + 1] = -30.0; xnom[8 + k * 10 + 2] = m0 * g * 1.5;
      state(k + 1, &xnom[8 + k * 10 + 3]);
    }
    const double pnom[9] = {g, alpha, r0[0], r0[1], r0[2], v0[0], v0[1], v0[2], m0};
    Rng r(3);
    dump_case(H, "rocket_hess", 61, [&](int j, long long, long long k) {
      if (j == 0) return xnom[k] + r.u(-1e-2, 1e-2);
      if (j == 1) return pnom[k] * r.u(0.8, 1.2);
      if (j == 2) return 1.0;
      return r.u(-1, 1);
    });
  }
  // config 3 --------------------------------------------------------------------------------------
  {
    Function leaf = mc_leaf();
    dump_tape(leaf, "mcstep");
    Rng r0(40);
    dump_case(leaf, "mcstep", 500, [&](int j, long long, long long) { return j == 0 ? r0.u(-1, 1) : r0.n(); });
    Function g = mc_rollout(100);
    dump_tape(g, "mc");
    Rng r(4);
    auto fill = [&](int j, long long, long long) { return j == 0 ? r.u(-1, 1) : 0.3 * r.n(); };
    dump_case(g, "mc", 200, fill);
    // reduce_out over both outputs: reference sums instances in index order (repmat.cpp:127-135)
    const long long N = 200;
    Function G = g.map("mcsum", "serial", N, std::vector<casadi_int>{}, std::vector<casadi_int>{0, 1});
    std::vector<double> x0(N * 4), W(N * 200), xs(4), Js(1);
    // same fill order as dump_case (input 0 for all instances, then input 1)
    {
      Rng r3(4);
      for (long long i = 0; i < N; ++i) for (int k = 0; k < 4; ++k) x0[i * 4 + k] = r3.u(-1, 1);
      for (long long i = 0; i < N; ++i) for (int k = 0; k < 200; ++k) W[i * 200 + k] = 0.3 * r3.n();
    }
    std::vector<const double*> arg(G.sz_arg(), nullptr);
    std::vector<double*> res(G.sz_res(), nullptr);
    std::vector<casadi_int> iw(G.sz_iw());
    std::vector<double> w(G.sz_w());
    arg[0] = x0.data(); arg[1] = W.data(); res[0] = xs.data(); res[1] = Js.data();
    casadi_assert(G(arg.data(), res.data(), iw.data(), w.data(), 0) == 0, "mcsum failed");
    FILE* fp = fopen((g_outdir + "/mc_sum.case").c_str(), "wb");
    wr(fp, "CCUCASE1", 8);
    wr64(fp, 1); wr64(fp, 0); wr64(fp, 2); wr64(fp, 4); wr64(fp, 1);
    wr(fp, xs.data(), 32); wr(fp, Js.data(), 8);
    fclose(fp);
  }
  // reference test function (function.py:658-696) ----------------------------------------------------
  {
    Function f = map_node_fun();
    dump_tape(f, "mapnode");
    Rng r(5);
    dump_case(f, "mapnode", 50, [&](int, long long, long long) { return r.u(0.1, 1.0); });
  }
  // operator coverage ---------------------------------------------------------------------------------
  {
    Function f = opcover();
    dump_tape(f, "opcover");
    Rng r(6);
    dump_case(f, "opcover", 4096, [&](int j, long long i, long long) {
      if (j == 2) return r.u(-0.999, 0.999);
      double v = r.u(-2, 2);
      if (i % 37 == 5) return 0.0;
      if (i % 37 == 6 && j == 0) return -0.0;
      return v;
    });
    // special values: every ordered pair from the list for (a,b); c cycles
    const double sp[] = {0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -3.0, inf, -inf, nan, 1e-310, -1e-310, 1e308, -1e308,
                         4.9e-324, 1e-17, 3.0, 7.5, -7.5, 710.0, -745.2, 1e22, 0.9999999999999999, 26.0, -26.0};
    const int ns = sizeof(sp) / sizeof(sp[0]);
    const double cs[] = {0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 0.999, -0.999, nan, 2.0, 0.7, -0.7, 0.71, -0.71, 1e-300};
    const int nc = sizeof(cs) / sizeof(cs[0]);
    dump_case(f, "opcover", ns * ns, [&](int j, long long i, long long) {
      if (j == 0) return sp[i / ns];
      if (j == 1) return sp[i % ns];
      return cs[i % nc];
    }, "opcover_special");
  }
  // OP_LIFT (calculus.hpp:1006-1010; f = x, the second operand is the initial guess of a lifted variable) and
  // OP_ASSIGN (calculus.hpp:600-606; f = x): neither is produced by expand() (MX lift expands to its first operand,
  // mx_function.cpp:1446), so the nodes are made with the public SX::binary / SX::unary constructors
  {
    SX a = SX::sym("a"), b = SX::sym("b");
    SX l1 = SX::binary(OP_LIFT, a * b + sin(a), b);
    SX l2 = SX::binary(OP_LIFT, b, a);
    SX as = SX::unary(OP_ASSIGN, a - b);
    Function fl("liftfun", {a, b}, {l1 * 3 + l2, SX::binary(OP_LIFT, as, a * b), SX::unary(OP_ASSIGN, l2) * as},
                {"a", "b"}, {"y", "z", "q"});
    dump_tape(fl, "liftfun");
    Rng r(11);
    dump_case(fl, "liftfun", 64, [&](int, long long k, long long) {
      if (k % 16 == 3) return 0.0;
      if (k % 16 == 7) return -0.0;
      if (k == 20) return std::numeric_limits<double>::infinity();
      if (k == 40) return std::numeric_limits<double>::quiet_NaN();
      return r.u(-2, 2);
    });
  }
  // config 4: KKT systems -- MX function [x = solve(K,b,solver); r = K*x-b] mapped serially, plus the symbolic
  // factorisation data the Linsol plugins compute in init (linsol_ldl.cpp:67-100, linsol_qr.cpp:67-84)
  {
    Sparsity sp = kkt_sparsity();
    std::vector<casadi_int> p, prinv, pc;
    Sparsity lt = sp.ldl(p, true);
    Sparsity spv, spr;
    sp.qr_sparse(spv, spr, prinv, pc, true);
    auto dumpv = [&](FILE* fp, const std::vector<casadi_int>& v) { wr64(fp, v.size()); wr(fp, v.data(), 8 * v.size()); };
    FILE* fp = fopen((g_outdir + "/kkt.sym").c_str(), "wb");
    wr(fp, "CCUSYM01", 8);
    dumpv(fp, sp.compress()); dumpv(fp, lt.compress()); dumpv(fp, p);
    dumpv(fp, spv.compress()); dumpv(fp, spr.compress()); dumpv(fp, prinv); dumpv(fp, pc);
    fclose(fp);
    for (std::string solver : {"ldl", "qr"}) {
      Function f = kkt_solve(solver);
      const long long N = 150;
      Rng r(7);
      dump_case(f, "kkt_" + solver, N, [&](int j, long long i, long long k) {
        if (j == 0) { static std::vector<double> v; static long long cur = -1; if (cur != i) { v = kkt_values(sp, i); cur = i; } return v[k]; }
        return r.u(-1, 1);
      });
    }
  }
  return 0;
}
