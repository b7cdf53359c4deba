// Tape builder and the traced runtime algorithms (see tape_builder.hpp).
//
// Every routine below walks the reference algorithm's loops over the SHARED sparsity pattern and records the
// floating-point operations in the reference's order, including the ones on structural zeros (0 - a*b is
// not -(a*b) for signed zeros, so nothing is simplified away).  Plain copies move handles and record nothing.
#include "tape_builder.hpp"

#include <algorithm>

namespace ccu {
namespace {
// enum Operation values used here (casadi/core/calculus.hpp:60-218)
enum { R_ADD = 1, R_SUB = 2, R_MUL = 3, R_DIV = 4, R_NEG = 5, R_SQRT = 10, R_LT = 19, R_LE = 20, R_EQ = 21, R_NOT = 23,
       R_FABS = 29, R_COPYSIGN = 31, R_IF_ELSE_ZERO = 32, R_CONST = 44, R_INPUT = 45, R_OUTPUT = 46 };
}  // namespace

TapeBuilder::V TapeBuilder::emit(int op, long long i0, long long i1, long long i2, double d) {
  op_.push_back(op);
  i0_.push_back(static_cast<int>(i0));
  i1_.push_back(static_cast<int>(i1));
  i2_.push_back(static_cast<int>(i2));
  d_.push_back(d);
  return i0;
}

TapeBuilder::V TapeBuilder::constant(double c) { return emit(R_CONST, next_++, 0, 0, c); }
TapeBuilder::V TapeBuilder::input(long long idx, long long nz) { return emit(R_INPUT, next_++, idx, nz, 0); }
void TapeBuilder::output(long long idx, long long nz, V v) { emit(R_OUTPUT, idx, v, nz, 0); }
TapeBuilder::V TapeBuilder::op(int refop, V a, V b) { return emit(refop, next_++, a, b < 0 ? a : b, 0); }

TapeBuilder::V TapeBuilder::select(V c, V a, V b) {
  // t = c ? a : +0, u = c ? +0 : b: exactly one is "active".  t+u loses only the sign of an active -0, which is
  // restored from sign(t)*sign(u) (the inactive one contributes +1).
  const V t = op(R_IF_ELSE_ZERO, c, a);
  const V u = op(R_IF_ELSE_ZERO, op(R_NOT, c), b);
  const V one = constant(1.0);
  const V sg = op(R_MUL, op(R_COPYSIGN, one, t), op(R_COPYSIGN, one, u));
  return op(R_COPYSIGN, op(R_ADD, t, u), sg);
}

// ---------------------------------------------------------------------------------------------------- LDL
void TapeBuilder::ldl(const long long* sp_a, const V* a, const long long* sp_lt, std::vector<V>* lt_, std::vector<V>* d_out,
                      const long long* p) {
  const long long n = sp_lt[1];
  const long long *lc = sp_lt + 2, *lr = sp_lt + 2 + n + 1;
  const long long *ac = sp_a + 2, *ar = sp_a + 2 + n + 1;
  const V zero = constant(0.0);
  std::vector<V> w(n, zero);
  std::vector<V>& lt = *lt_;
  std::vector<V>& d = *d_out;
  lt.assign(lc[n], zero);
  d.assign(n, zero);
  // A(p,p) scattered into the pattern of L^T and into D (casadi_ldl.hpp:34-43)
  for (long long c = 0; c < n; ++c) {
    const long long cp = p[c];
    for (long long k = ac[cp]; k < ac[cp + 1]; ++k) w[ar[k]] = a[k];
    for (long long k = lc[c]; k < lc[c + 1]; ++k) lt[k] = w[p[lr[k]]];
    d[c] = w[p[c]];
    for (long long k = ac[cp]; k < ac[cp + 1]; ++k) w[ar[k]] = zero;
  }
  // up-looking elimination (casadi_ldl.hpp:44-58)
  for (long long c = 0; c < n; ++c) {
    for (long long k = lc[c]; k < lc[c + 1]; ++k) {
      const long long r = lr[k];
      for (long long k2 = lc[r]; k2 < lc[r + 1]; ++k2) lt[k] = op(R_SUB, lt[k], op(R_MUL, lt[k2], w[lr[k2]]));
      w[r] = lt[k];
      lt[k] = op(R_DIV, lt[k], d[r]);
      d[c] = op(R_SUB, d[c], op(R_MUL, w[r], lt[k]));
    }
    for (long long k = lc[c]; k < lc[c + 1]; ++k) w[lr[k]] = zero;
  }
}

void TapeBuilder::ldl_solve(V* x, long long nrhs, const long long* sp_lt, const V* lt, const V* d, const long long* p) {
  const long long n = sp_lt[1];
  const long long *ci = sp_lt + 2, *ri = sp_lt + 2 + n + 1;
  std::vector<V> w(n);
  for (long long rhs = 0; rhs < nrhs; ++rhs, x += n) {
    for (long long i = 0; i < n; ++i) w[i] = x[p[i]];
    // L w = w: L^T is stored, so its transpose is applied column by column (casadi_ldl.hpp:64-85, tr = true)
    for (long long c = 0; c < n; ++c)
      for (long long k = ci[c]; k < ci[c + 1]; ++k) w[c] = op(R_SUB, w[c], op(R_MUL, lt[k], w[ri[k]]));
    for (long long i = 0; i < n; ++i) w[i] = op(R_DIV, w[i], d[i]);
    for (long long c = n - 1; c >= 0; --c)
      for (long long k = ci[c + 1] - 1; k >= ci[c]; --k) w[ri[k]] = op(R_SUB, w[ri[k]], op(R_MUL, lt[k], w[c]));
    for (long long i = 0; i < n; ++i) x[p[i]] = w[i];
  }
}

TapeBuilder::V TapeBuilder::ldl_zero_pivots(const V* d, long long n) {
  const V zero = constant(0.0);
  V cnt = zero;
  for (long long c = 0; c < n; ++c) cnt = op(R_ADD, cnt, op(R_EQ, d[c], zero));
  return cnt;
}

// ----------------------------------------------------------------------------------------------------- QR
void TapeBuilder::qr(const long long* sp_a, const V* a, const long long* sp_v, std::vector<V>* v_, const long long* sp_r,
                     std::vector<V>* r_, std::vector<V>* beta_, const long long* prinv, const long long* pc) {
  const long long ncol = sp_a[1], nrow = sp_v[0];
  const long long *ac = sp_a + 2, *ar = sp_a + 2 + ncol + 1;
  const long long *vc = sp_v + 2, *vr = sp_v + 2 + ncol + 1;
  const long long *rc = sp_r + 2, *rr = sp_r + 2 + ncol + 1;
  const V zero = constant(0.0);
  std::vector<V> x(nrow, zero);
  std::vector<V>& v = *v_;
  std::vector<V>& rv = *r_;
  std::vector<V>& beta = *beta_;
  v.assign(vc[ncol], zero);
  rv.assign(rc[ncol], zero);
  beta.assign(ncol, zero);
  long long nr = 0;  // R is filled sequentially (casadi_qr.hpp: *nz_r++ = ...)
  for (long long c = 0; c < ncol; ++c) {
    for (long long k = ac[pc[c]]; k < ac[pc[c] + 1]; ++k) x[prinv[ar[k]]] = a[k];
    // apply the previous reflections that touch this column: strictly upper part of R
    for (long long k = rc[c]; k < rc[c + 1] && rr[k] < c; ++k) {
      const long long r = rr[k];
      V alpha = zero;
      for (long long k1 = vc[r]; k1 < vc[r + 1]; ++k1) alpha = op(R_ADD, alpha, op(R_MUL, v[k1], x[vr[k1]]));
      alpha = op(R_MUL, alpha, beta[r]);
      for (long long k1 = vc[r]; k1 < vc[r + 1]; ++k1) x[vr[k1]] = op(R_SUB, x[vr[k1]], op(R_MUL, alpha, v[k1]));
      rv[nr++] = x[r];
      x[r] = zero;
    }
    for (long long k = vc[c]; k < vc[c + 1]; ++k) {
      v[k] = x[vr[k]];
      x[vr[k]] = zero;
    }
    // casadi_house (casadi_qr.hpp:24-41) on column c of V
    {
      V* hv = v.data() + vc[c];
      const long long nv = vc[c + 1] - vc[c];
      const V v0 = hv[0];
      V sigma = zero;
      for (long long i = 1; i < nv; ++i) sigma = op(R_ADD, sigma, op(R_MUL, hv[i], hv[i]));
      const V s = op(R_SQRT, op(R_ADD, op(R_MUL, v0, v0), sigma));
      const V sigma_is_zero = op(R_EQ, sigma, zero);
      const V v0_nonpos = op(R_LE, v0, zero);
      const V inner = select(v0_nonpos, op(R_SUB, v0, s), op(R_DIV, op(R_NEG, sigma), op(R_ADD, v0, s)));
      hv[0] = select(sigma_is_zero, constant(1.0), inner);
      beta[c] = select(sigma_is_zero, op(R_MUL, constant(2.0), v0_nonpos), op(R_DIV, constant(-1.0), op(R_MUL, s, hv[0])));
      rv[nr++] = s;
    }
  }
}

namespace {
// x = Q*x (tr = false) or Q'*x (tr = true), Q given by the Householder vectors (casadi_qr.hpp:104-124)
void qr_mv(TapeBuilder& B, const long long* sp_v, const TapeBuilder::V* v, const TapeBuilder::V* beta, TapeBuilder::V* x,
           bool tr, TapeBuilder::V zero) {
  const long long ncol = sp_v[1];
  const long long *ci = sp_v + 2, *ri = sp_v + 2 + ncol + 1;
  for (long long c1 = 0; c1 < ncol; ++c1) {
    const long long c = tr ? c1 : ncol - 1 - c1;
    TapeBuilder::V alpha = zero;
    for (long long k = ci[c]; k < ci[c + 1]; ++k) alpha = B.op(R_ADD, alpha, B.op(R_MUL, v[k], x[ri[k]]));
    alpha = B.op(R_MUL, alpha, beta[c]);
    for (long long k = ci[c]; k < ci[c + 1]; ++k) x[ri[k]] = B.op(R_SUB, x[ri[k]], B.op(R_MUL, alpha, v[k]));
  }
}

// R x = b / R' x = b by substitution (casadi_qr.hpp:129-161)
void qr_trs(TapeBuilder& B, const long long* sp_r, const TapeBuilder::V* nz, TapeBuilder::V* x, bool tr) {
  const long long ncol = sp_r[1];
  const long long *ci = sp_r + 2, *ri = sp_r + 2 + ncol + 1;
  if (tr) {
    for (long long c = 0; c < ncol; ++c)
      for (long long k = ci[c]; k < ci[c + 1]; ++k) {
        const long long r = ri[k];
        x[c] = r == c ? B.op(R_DIV, x[c], nz[k]) : B.op(R_SUB, x[c], B.op(R_MUL, nz[k], x[r]));
      }
  } else {
    for (long long c = ncol - 1; c >= 0; --c)
      for (long long k = ci[c + 1] - 1; k >= ci[c]; --k) {
        const long long r = ri[k];
        x[r] = r == c ? B.op(R_DIV, x[r], nz[k]) : B.op(R_SUB, x[r], B.op(R_MUL, nz[k], x[c]));
      }
  }
}
}  // namespace

void TapeBuilder::qr_solve(V* x, long long nrhs, bool tr, const long long* sp_v, const V* v, const long long* sp_r,
                           const V* r, const V* beta, const long long* prinv, const long long* pc) {
  const long long nrow_ext = sp_v[0], ncol = sp_v[1];
  const V zero = constant(0.0);
  std::vector<V> w(std::max(nrow_ext, ncol), zero);
  for (long long rhs = 0; rhs < nrhs; ++rhs, x += ncol) {
    if (tr) {
      for (long long c = 0; c < ncol; ++c) w[c] = x[pc[c]];
      qr_trs(*this, sp_r, r, w.data(), true);
      qr_mv(*this, sp_v, v, beta, w.data(), false, zero);
      for (long long c = 0; c < ncol; ++c) x[c] = w[prinv[c]];
    } else {
      for (long long c = 0; c < nrow_ext; ++c) w[c] = zero;
      for (long long c = 0; c < ncol; ++c) w[prinv[c]] = x[c];
      qr_mv(*this, sp_v, v, beta, w.data(), true, zero);
      qr_trs(*this, sp_r, r, w.data(), false);
      for (long long c = 0; c < ncol; ++c) x[pc[c]] = w[c];
    }
  }
}

TapeBuilder::V TapeBuilder::qr_nullity(const V* r, const long long* sp_r, double eps) {
  const long long ncol = sp_r[1];
  const long long* rc = sp_r + 2;
  const V zero = constant(0.0), e = constant(eps);
  V cnt = zero;
  for (long long c = 0; c < ncol; ++c) cnt = op(R_ADD, cnt, op(R_LT, op(R_FABS, r[rc[c + 1] - 1]), e));
  return cnt;
}

// ------------------------------------------------------------------------------------------------- mtimes
void TapeBuilder::mtimes(const V* x, const long long* sp_x, const V* y, const long long* sp_y, V* z, const long long* sp_z) {
  const long long ncol_x = sp_x[1], ncol_y = sp_y[1], ncol_z = sp_z[1];
  const long long *cx = sp_x + 2, *rx = sp_x + 2 + ncol_x + 1;
  const long long *cy = sp_y + 2, *ry = sp_y + 2 + ncol_y + 1;
  const long long *cz = sp_z + 2, *rz = sp_z + 2 + ncol_z + 1;
  // dense work column; rows outside z's pattern accumulate values the reference never reads back either
  std::vector<V> w(sp_x[0], constant(0.0));
  for (long long cc = 0; cc < ncol_y; ++cc) {
    for (long long kk = cz[cc]; kk < cz[cc + 1]; ++kk) w[rz[kk]] = z[kk];
    for (long long kk = cy[cc]; kk < cy[cc + 1]; ++kk) {
      const long long rr = ry[kk];
      for (long long kk1 = cx[rr]; kk1 < cx[rr + 1]; ++kk1) w[rx[kk1]] = op(R_ADD, w[rx[kk1]], op(R_MUL, x[kk1], y[kk]));
    }
    for (long long kk = cz[cc]; kk < cz[cc + 1]; ++kk) z[kk] = w[rz[kk]];
  }
}

TapeSource TapeBuilder::source(const std::vector<long long>& nnz_in, const std::vector<long long>& nnz_out) const {
  TapeSource s;
  s.n_instr = static_cast<long long>(op_.size());
  s.op = op_.data(); s.i0 = i0_.data(); s.i1 = i1_.data(); s.i2 = i2_.data(); s.d = d_.data();
  s.sz_w = next_;
  s.nnz_in = nnz_in;
  s.nnz_out = nnz_out;
  return s;
}

}  // namespace ccu
