// Tape builder: records scalar operations into the SXFunction tape format (ScalarAtomic stream,
// casadi/core/sx_function.hpp:37-44) so that everything the device evaluates is ONE kind of program.
//
// Why it exists: the Linsol calls inside mapped functions (LinsolCall::eval, casadi/core/solve_impl.hpp:57-73
// -> casadi_ldl / casadi_qr, casadi/core/runtime/casadi_ldl.hpp, casadi_qr.hpp) run with a sparsity pattern
// that is SHARED by the whole batch.  With the pattern and the permutations fixed, the sequence of
// floating-point operations of the factorisation and of the triangular solves is data-independent: a
// straight-line program.  The builder executes those algorithms once, symbolically, over value handles and
// records every +,-,*,/,sqrt in the reference's order; the resulting tape is evaluated per instance by the
// same interpreter / specialised kernels as any SX tape (K3/K4 of SURVEY 2 = traced tapes on K1).  The MX
// glue around such calls (casadi_mtimes, element-wise operations) is traced the same way, which is how
// CudaMap puts a non-expandable MX function (BASELINE config 5) on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tape_compile.hpp"

namespace ccu {

class TapeBuilder {
 public:
  typedef long long V;  // value handle = work slot of the recorded tape (single assignment)

  V constant(double c);
  V input(long long idx, long long nz);
  void output(long long idx, long long nz, V v);
  // refop: enum Operation numbering (calculus.hpp:60-218); b is ignored for unary operations
  V op(int refop, V a, V b = -1);
  // c != 0 ? a : b with the selected operand's bits preserved (if_else on doubles, calculus.hpp:296)
  V select(V c, V a, V b);

  // ---- traced runtime algorithms; patterns are compressed CCS [nrow, ncol, colind[ncol+1], row[nnz]] --------
  // casadi_ldl (casadi_ldl.hpp:25-59): a = nnz(A) values -> lt (nnz(L^T)), d (n)
  void ldl(const long long* sp_a, const V* a, const long long* sp_lt, std::vector<V>* lt, std::vector<V>* d,
           const long long* p);
  // casadi_ldl_solve (casadi_ldl.hpp:90-109): x (n*nrhs) in place
  void ldl_solve(V* x, long long nrhs, const long long* sp_lt, const V* lt, const V* d, const long long* p);
  // casadi_qr (casadi_qr.hpp:53-95) incl. casadi_house (:24-41)
  void qr(const long long* sp_a, const V* a, const long long* sp_v, std::vector<V>* v, const long long* sp_r,
          std::vector<V>* r, std::vector<V>* beta, const long long* prinv, const long long* pc);
  // casadi_qr_solve (casadi_qr.hpp:167-197)
  void qr_solve(V* x, long long nrhs, bool tr, const long long* sp_v, const V* v, const long long* sp_r, const V* r,
                const V* beta, const long long* prinv, const long long* pc);
  // number of |R_cc| < eps (casadi_qr_singular, casadi_qr.hpp:202-227) as a value
  V qr_nullity(const V* r, const long long* sp_r, double eps);
  // number of zero pivots in D (the condition LinsolLdl::nfact warns about, linsol_ldl.cpp:122-124)
  V ldl_zero_pivots(const V* d, long long n);
  // casadi_mtimes, tr = false (casadi_mtimes.hpp:22-75): z += x*y
  void mtimes(const V* x, const long long* sp_x, const V* y, const long long* sp_y, V* z, const long long* sp_z);

  long long n_values() const { return next_; }
  // view of the recorded tape (valid until the next recording call)
  TapeSource source(const std::vector<long long>& nnz_in, const std::vector<long long>& nnz_out) const;
  const std::string& error() const { return err_; }

 private:
  std::vector<int> op_, i0_, i1_, i2_;
  std::vector<double> d_;
  long long next_ = 0;
  std::string err_;
  V emit(int op, long long i0, long long i1, long long i2, double d);
};

}  // namespace ccu
