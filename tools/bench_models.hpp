// Benchmark / parity models for the five BASELINE.json configs, written against the
// reference's public C++ API (casadi/casadi.hpp).  TEST / MEASUREMENT INFRASTRUCTURE: model definitions only (no
// arithmetic of the hot path); compiled into oracle/_ref/bin/* (fixture generator, reference-arm bench), the
// integration test and tools/cuda_bench (the end-to-end benchmark of f.map(N,"cuda") through the patched reference).
// Model definitions follow SURVEY.md Appendix A.
#pragma once
#include <casadi/casadi.hpp>
#include <algorithm>
#include <random>
#include <string>
#include <vector>

namespace ccu_models {
using namespace casadi;

// ---- config 0: cart-pole, 4-state ODE, M RK4 substeps of h=0.01 inlined ---------------------------
inline SX cartpole_ode(const SX& x, const SX& u) {
  const double mc = 1.0, mp = 0.1, l = 0.5, g = 9.81;
  SX th = x(2), om = x(3);
  SX s = sin(th), c = cos(th);
  SX tmp = (u + mp * l * om * om * s) / (mc + mp);
  SX thdd = (g * s - c * tmp) / (l * (4.0 / 3.0 - mp * c * c / (mc + mp)));
  SX xdd = tmp - mp * l * thdd * c / (mc + mp);
  return vertcat(x(1), xdd, om, thdd);
}

inline Function cartpole(int M = 4) {
  SX x = SX::sym("x", 4), u = SX::sym("u");
  const double h = 0.01;
  SX xk = x;
  for (int i = 0; i < M; ++i) {
    SX k1 = cartpole_ode(xk, u);
    SX k2 = cartpole_ode(xk + h / 2 * k1, u);
    SX k3 = cartpole_ode(xk + h / 2 * k2, u);
    SX k4 = cartpole_ode(xk + h * k3, u);
    xk = xk + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4);
  }
  return Function("cartpole", {x, u}, {xk}, {"x", "u"}, {"xf"});
}

// ---- config 1: quadrotor, 12 states, 4 rotor thrusts, nsteps RK4 steps of h=0.005 ------------------
inline SX quad_ode(const SX& x, const SX& u) {
  const double m = 1.2, g = 9.81, Ix = 0.02, Iy = 0.02, Iz = 0.04, l = 0.25, kq = 0.02;
  SX phi = x(6), th = x(7), psi = x(8), p = x(9), q = x(10), r = x(11);
  SX T = u(0) + u(1) + u(2) + u(3);
  SX tx = l * (u(1) - u(3)), ty = l * (u(2) - u(0)), tz = kq * (u(0) - u(1) + u(2) - u(3));
  SX cph = cos(phi), sph = sin(phi), cth = cos(th), sth = sin(th), cps = cos(psi), sps = sin(psi);
  SX ax = (cph * sth * cps + sph * sps) * T / m;
  SX ay = (cph * sth * sps - sph * cps) * T / m;
  SX az = cph * cth * T / m - g;
  SX tth = sth / cth;
  SX phid = p + (q * sph + r * cph) * tth;
  SX thd = q * cph - r * sph;
  SX psid = (q * sph + r * cph) / cth;
  SX pd = (tx - (Iz - Iy) * q * r) / Ix;
  SX qd = (ty - (Ix - Iz) * p * r) / Iy;
  SX rd = (tz - (Iy - Ix) * p * q) / Iz;
  return vertcat(std::vector<SX>{x(3), x(4), x(5), ax, ay, az, phid, thd, psid, pd, qd, rd});
}

inline Function quadrotor(int nsteps = 20) {
  SX x = SX::sym("x", 12), u = SX::sym("u", 4);
  const double h = 0.005;
  SX xk = x;
  for (int i = 0; i < nsteps; ++i) {
    SX k1 = quad_ode(xk, u);
    SX k2 = quad_ode(xk + h / 2 * k1, u);
    SX k3 = quad_ode(xk + h / 2 * k2, u);
    SX k4 = quad_ode(xk + h * k3, u);
    xk = xk + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4);
  }
  return Function("quad", {x, u}, {xk}, {"x", "u"}, {"xf"});
}

// ---- config 2: rocket-landing OCP NLP; hess_lag via Function::factory ------------------------------
inline SX rocket_ode(const SX& s, const SX& c, const SX& g, const SX& alpha) {
  SX m = s(6);
  SX tn = sqrt(dot(c, c) + 1e-8);
  return vertcat(std::vector<SX>{s(3), s(4), s(5), c(0) / m, c(1) / m, c(2) / m - g, -alpha * tn});
}

inline Function rocket_nlp(int K = 20) {
  // decision vector: [tf, s0(7), (c_k(3), s_{k+1}(7)) k<K]
  int nx = 1 + 7 + K * 10;
  SX X = SX::sym("x", nx);
  SX P = SX::sym("p", 9);  // g, alpha, r0(3), v0(3), m0
  SX g = P(0), alpha = P(1);
  SX tf = X(0);
  SX h = tf / K;
  auto S = [&](int k) { return k == 0 ? X(Slice(1, 8)) : X(Slice(8 + (k - 1) * 10 + 3, 8 + (k - 1) * 10 + 10)); };
  auto C = [&](int k) { return X(Slice(8 + k * 10, 8 + k * 10 + 3)); };
  SX f = 0;
  std::vector<SX> gv;
  gv.push_back(S(0) - P(Slice(2, 9)));
  for (int k = 0; k < K; ++k) {
    SX s = S(k), c = C(k);
    SX k1 = rocket_ode(s, c, g, alpha);
    SX k2 = rocket_ode(s + h / 2 * k1, c, g, alpha);
    SX k3 = rocket_ode(s + h / 2 * k2, c, g, alpha);
    SX k4 = rocket_ode(s + h * k3, c, g, alpha);
    SX sn = s + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4);
    gv.push_back(S(k + 1) - sn);
    gv.push_back(dot(c, c));
    f += h * sqrt(dot(c, c) + 1e-8);
  }
  gv.push_back(S(K)(Slice(0, 6)));
  return Function("rocket_nlp", {X, P}, {f, vertcat(gv)}, {"x", "p"}, {"f", "g"});
}

inline Function rocket_hess_lag(int K = 20) {
  Function nlp = rocket_nlp(K);
  // same factory call as casadi/solvers/sqpmethod.cpp:280-284
  return nlp.factory("hess_lag", {"x", "p", "lam:f", "lam:g"}, {"hess:gamma:x:x"}, {{"gamma", {"f", "g"}}});
}

// ---- config 3: mapaccum Monte-Carlo time stepping ---------------------------------------------------
// leaf: damped Duffing-type pair of oscillators driven by noise w, stage cost x'x + 0.1 w'w
inline Function mc_leaf() {
  SX x = SX::sym("x", 4), w = SX::sym("w", 2);
  const double h = 0.02;
  SX q1 = x(0), v1 = x(1), q2 = x(2), v2 = x(3);
  SX a1 = -q1 - 0.3 * q1 * q1 * q1 - 0.1 * v1 + 0.5 * (q2 - q1) + w(0);
  SX a2 = -sin(q2) - 0.1 * v2 + 0.5 * (q1 - q2) + w(1);
  SX xn = vertcat(std::vector<SX>{q1 + h * v1, v1 + h * a1, q2 + h * v2, v2 + h * a2});
  SX cost = dot(x, x) + 0.1 * dot(w, w);
  return Function("mcstep", {x, w}, {xn, cost}, {"x", "w"}, {"xn", "cost"});
}

// mc(x0, W[2xT]) -> (xT, sum_t cost_t): f.mapaccum(T) (function.cpp:668-747) then expanded to one SX tape
inline Function mc_rollout(int T = 100) {
  Function f = mc_leaf();
  Function acc = f.mapaccum(T);
  MX x0 = MX::sym("x0", 4), W = MX::sym("W", 2, T);
  std::vector<MX> r = acc(std::vector<MX>{x0, W});
  MX xT = r[0](Slice(), T - 1);
  MX J = sum2(r[1]);
  Function g("mc", {x0, W}, {xT, J}, {"x0", "W"}, {"xT", "J"});
  return g.expand();
}

// ---- config 4: KKT system  K=[[H,A'],[A,D]], n=60 ---------------------------------------------------
inline Sparsity kkt_sparsity() {
  Sparsity H = Sparsity::banded(40, 3);
  std::vector<casadi_int> ar, ac;
  for (int i = 0; i < 20; ++i) { ar.push_back(i); ac.push_back(2 * i); ar.push_back(i); ac.push_back(2 * i + 1); }
  Sparsity A = Sparsity::triplet(20, 40, ar, ac);
  Sparsity D = Sparsity::diag(20);
  return blockcat(std::vector<std::vector<Sparsity>>{{H, A.T()}, {A, D}});
}

// sf(K,b) -> (x = solve(K,b,"ldl"|"qr"), r = K*x-b)   MXFunction, cannot be expand()ed
inline Function kkt_solve(const std::string& solver = "ldl") {
  Sparsity sp = kkt_sparsity();
  MX K = MX::sym("K", sp), b = MX::sym("b", 60);
  MX x = solve(K, b, solver);
  MX r = mtimes(K, x) - b;
  return Function("kkt_" + solver, {K, b}, {x, r}, {"K", "b"}, {"x", "r"});
}

// fill one KKT instance (values per SURVEY Appendix A), returns nnz values in CCS order
inline std::vector<double> kkt_values(const Sparsity& sp, casadi_int inst) {
  std::vector<double> v(sp.nnz());
  const casadi_int* colind = sp.colind();
  const casadi_int* row = sp.row();
  double scale = 1 + 1e-3 * static_cast<double>(inst % 97);
  for (casadi_int c = 0; c < sp.size2(); ++c)
    for (casadi_int k = colind[c]; k < colind[c + 1]; ++k) {
      casadi_int r = row[k];
      double val = r == c ? (r < 40 ? 10.0 : -1e-2) : 0.3 + 0.01 * static_cast<double>((7 * std::min(r, c) + 3 * std::max(r, c)) % 11);
      v[k] = val * scale;
    }
  return v;
}

// ---- reference test function: test/python/function.py:658-696 (test_map_node) ----------------------
inline Function map_node_fun() {
  SX x = SX::sym("x"), y = SX::sym("y", 2), z = SX::sym("z", 2, 2), v = SX::sym("v", Sparsity::upper(3));
  return Function("f", {x, y, z, v}, {mtimes(z, y) + x, sin(y * x).T(), v / x});
}

// ---- operator coverage: every scalar-evaluable opcode of calculus.hpp:1302-1355 (except PRINTME) ---
inline Function opcover() {
  SX a = SX::sym("a"), b = SX::sym("b"), c = SX::sym("c");
  std::vector<SX> o;
  o.push_back(a + b); o.push_back(a - b); o.push_back(a * b); o.push_back(a / b); o.push_back(-a);
  o.push_back(exp(a)); o.push_back(log(fabs(a) + 0.1)); o.push_back(log(a));
  o.push_back(pow(fabs(a) + 0.5, b)); o.push_back(pow(a, 3.0)); o.push_back(pow(a, b));
  o.push_back(constpow(fabs(a), SX(2.5)));
  o.push_back(sqrt(fabs(a))); o.push_back(sqrt(a)); o.push_back(sq(a)); o.push_back(2 * a);
  o.push_back(sin(a)); o.push_back(cos(a)); o.push_back(tan(a));
  o.push_back(asin(c)); o.push_back(acos(c)); o.push_back(atan(a));
  o.push_back(a < b); o.push_back(a <= b); o.push_back(a == b); o.push_back(a != b);
  o.push_back(!a); o.push_back(a && b); o.push_back(a || b);
  o.push_back(floor(a)); o.push_back(ceil(a)); o.push_back(fmod(a, b)); o.push_back(remainder(a, b));
  o.push_back(fabs(a)); o.push_back(sign(a)); o.push_back(copysign(a, b));
  o.push_back(if_else_zero(a > 0, b)); o.push_back(if_else(a < b, a * c, b - c));
  o.push_back(erf(a)); o.push_back(fmin(a, b)); o.push_back(fmax(a, b)); o.push_back(1 / a);
  o.push_back(sinh(a)); o.push_back(cosh(a)); o.push_back(tanh(a));
  o.push_back(asinh(a)); o.push_back(acosh(fabs(a) + 1)); o.push_back(acosh(a)); o.push_back(atanh(c));
  o.push_back(atan2(a, b)); o.push_back(erfinv(c)); o.push_back(erfinv(a));
  o.push_back(log1p(a)); o.push_back(expm1(a)); o.push_back(hypot(a, b));
  o.push_back(exp(5 * a)); o.push_back(sin(1e3 * a)); o.push_back(cos(1e6 * b)); o.push_back(tan(40 * a));
  o.push_back(log1p(c * 1e-9)); o.push_back(expm1(c * 1e-9)); o.push_back(pow(b, a * 7));
  return Function("opcover", {a, b, c}, {vertcat(o)}, {"a", "b", "c"}, {"y"});
}


// ---- synthetic inputs of SURVEY 8(d) for a workload `kind` (cartpole | quad | rocket | mc | kkt), AoS per instance ----
inline void bench_inputs(const Function& f, long long n, unsigned long long seed, const std::string& kind,
                         std::vector<std::vector<double>>& in, const std::vector<bool>& reduce_in = {}) {
  std::mt19937_64 g(seed);
  auto u = [&](double a, double b) { return a + (b - a) * std::generate_canonical<double, 53>(g); };
  const double hover = 1.2 * 9.81 / 4;
  Sparsity ksp = kkt_sparsity();
  in.resize(f.n_in());
  for (casadi_int k = 0; k < f.n_in(); ++k) {
    const long long nz = f.nnz_in(k);
    const long long cnt = (k < static_cast<casadi_int>(reduce_in.size()) && reduce_in[k]) ? 1 : n;
    in[k].resize(cnt * nz);
    for (long long i = 0; i < cnt; ++i) {
      if (kind == "kkt" && k == 0) {
        std::vector<double> v = kkt_values(ksp, i);
        std::copy(v.begin(), v.end(), in[k].begin() + i * nz);
        continue;
      }
      for (long long e = 0; e < nz; ++e) {
        double v;
        if (kind == "cartpole") v = k == 0 ? u(-0.5, 0.5) : u(-1, 1);
        else if (kind == "quad") v = k == 0 ? u(-0.3, 0.3) : k == 1 ? hover * (1 + u(-0.1, 0.1)) : u(-1, 1);
        else if (kind == "rocket") v = k == 0 ? 1.0 + u(-1e-2, 1e-2) + 0.01 * e : k == 1 ? 1.0 * u(0.8, 1.2) : k == 2 ? 1.0 : u(-1, 1);
        else if (kind == "mc") v = k == 0 ? u(-1, 1) : 0.3 * u(-1.7, 1.7);
        else v = u(-1, 1);
        in[k][i * nz + e] = v;
      }
    }
  }
}

// the functions of a benchmark workload name and the input kind they take
inline std::vector<Function> bench_workload(const std::string& wl, std::string* kind) {
  if (wl == "cartpole") { *kind = "cartpole"; return {cartpole(4)}; }
  if (wl == "quad") { *kind = "quad"; return {quadrotor(20)}; }
  if (wl == "quad_jac") { *kind = "quad"; return {quadrotor(20).jacobian()}; }
  if (wl == "quad_ms") { *kind = "quad"; Function F = quadrotor(20); return {F, F.jacobian()}; }
  if (wl == "rocket_hess") { *kind = "rocket"; return {rocket_hess_lag(20)}; }
  if (wl == "mc") { *kind = "mc"; return {mc_rollout(100)}; }
  if (wl == "kkt_ldl") { *kind = "kkt"; return {kkt_solve("ldl")}; }
  if (wl == "kkt_qr") { *kind = "kkt"; return {kkt_solve("qr")}; }
  casadi_error("unknown workload " + wl);
}

}  // namespace ccu_models
