/*
 * CudaMap -- see cuda_map.hpp.  New file for casadi/core/.
 */
#include "cuda_map.hpp"
#include "bspline.hpp"
#include "integrator_impl.hpp"
#include "linsol.hpp"
#include "mapsum.hpp"
#include "multiplication.hpp"
#include "rootfinder_impl.hpp"
#include <casadi/solvers/bspline_interpolant.hpp>  // S_ of the B-spline interpolant (layout only)
#include <casadi/solvers/linear_interpolant.hpp>  // data members of the lookup-table plugin (layout only; nothing is linked)
#include <casadi/solvers/fast_newton.hpp>  // option members of the FastNewton plugin class (layout only)
#include <casadi/solvers/newton.hpp>  // option members of the Newton plugin class (layout only; nothing is linked)
#include "mx_node.hpp"
#include "solve.hpp"
#include "switch.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <mutex>

namespace casadi {

  namespace {
    // The slice of include/casadi_cuda.h that CudaMap binds (plain C, resolved with dlsym)
    typedef long long ccu_int;
    struct CudaLib {
      void* handle = nullptr;
      std::string error;
      int (*abi_version)() = nullptr;
      const char* (*last_error)() = nullptr;
      int (*device_count)() = nullptr;
      // one tape replicated on the devices of this process (a single device is the common case)
      void* (*multi_create)(ccu_int, const int*, const int*, const int*, const int*, const double*, ccu_int,
                            ccu_int, const ccu_int*, ccu_int, const ccu_int*, int, const int*) = nullptr;
      void* (*builder_finish_multi)(void*, ccu_int, const ccu_int*, ccu_int, const ccu_int*, int, const int*) = nullptr;
      void (*multi_destroy)(void*) = nullptr;
      int (*multi_size)(const void*) = nullptr;
      void* (*multi_tape)(void*, int) = nullptr;
      int (*multi_eval_host)(void*, ccu_int, const double* const*, double* const*, const int*, const int*,
                             const int*, const int*) = nullptr;
      int (*last_eval_stats)(const void*, double*) = nullptr;
      // tape builder (MX functions that cannot be expanded are lowered to one scalar tape)
      void* (*builder_create)() = nullptr;
      void (*builder_destroy)(void*) = nullptr;
      ccu_int (*builder_const)(void*, double) = nullptr;
      ccu_int (*builder_input)(void*, ccu_int, ccu_int) = nullptr;
      ccu_int (*builder_op)(void*, int, ccu_int, ccu_int) = nullptr;
      int (*builder_output)(void*, ccu_int, ccu_int, ccu_int) = nullptr;
      int (*builder_ldl)(void*, const ccu_int*, const ccu_int*, const ccu_int*, const ccu_int*, ccu_int*, ccu_int,
                         ccu_int*) = nullptr;
      int (*builder_qr)(void*, const ccu_int*, const ccu_int*, const ccu_int*, const ccu_int*, const ccu_int*,
                        const ccu_int*, ccu_int*, ccu_int, int, double, ccu_int*) = nullptr;
      int (*builder_mtimes)(void*, const ccu_int*, const ccu_int*, const ccu_int*, const ccu_int*, ccu_int*,
                            const ccu_int*) = nullptr;
      // single-device entry points (the Newton driver keeps its state on the device between launches)
      void* (*tape_create)(ccu_int, const int*, const int*, const int*, const int*, const double*, ccu_int,
                           ccu_int, const ccu_int*, ccu_int, const ccu_int*, int) = nullptr;
      void (*tape_destroy)(void*) = nullptr;
      int (*eval_reduce_device)(void*, ccu_int, const double* const*, double* const*, const int*, const int*, int,
                                void*) = nullptr;
      int (*set_device)(int) = nullptr;
      void* (*dev_malloc)(ccu_int) = nullptr;
      int (*dev_free)(void*) = nullptr;
      int (*memcpy_h2d)(void*, const void*, ccu_int, void*) = nullptr;
      int (*memcpy_d2h)(void*, const void*, ccu_int, void*) = nullptr;
      int (*stream_sync)(void*) = nullptr;
      ccu_int (*builder_select)(void*, ccu_int, ccu_int, ccu_int) = nullptr;
      ccu_int (*builder_export)(const void*, int*, int*, int*, int*, double*, ccu_int, ccu_int*) = nullptr;
    };

    CudaLib& cuda_lib() {
      static CudaLib lib;
      static std::once_flag once;
      std::call_once(once, [] {
        const char* env = getenv("CASADI_CUDA_LIB");
        const std::string name = env ? env : "libcasadi_cuda.so";
        lib.handle = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib.handle) {
          lib.error = "Cannot load '" + name + "': " + dlerror()
            + " (set CASADI_CUDA_LIB or the library search path)";
          return;
        }
        bool ok = true;
        auto sym = [&](const char* s) { void* p = dlsym(lib.handle, s); if (!p) ok = false; return p; };
        lib.abi_version = reinterpret_cast<decltype(lib.abi_version)>(sym("ccu_abi_version"));
        lib.last_error = reinterpret_cast<decltype(lib.last_error)>(sym("ccu_last_error"));
        lib.device_count = reinterpret_cast<decltype(lib.device_count)>(sym("ccu_device_count"));
        lib.multi_create = reinterpret_cast<decltype(lib.multi_create)>(sym("ccu_multi_create"));
        lib.builder_finish_multi =
          reinterpret_cast<decltype(lib.builder_finish_multi)>(sym("ccu_builder_finish_multi"));
        lib.multi_destroy = reinterpret_cast<decltype(lib.multi_destroy)>(sym("ccu_multi_destroy"));
        lib.multi_size = reinterpret_cast<decltype(lib.multi_size)>(sym("ccu_multi_size"));
        lib.multi_tape = reinterpret_cast<decltype(lib.multi_tape)>(sym("ccu_multi_tape"));
        lib.multi_eval_host = reinterpret_cast<decltype(lib.multi_eval_host)>(sym("ccu_multi_eval_host_grouped"));
        lib.last_eval_stats = reinterpret_cast<decltype(lib.last_eval_stats)>(sym("ccu_tape_last_eval_stats"));
        lib.builder_create = reinterpret_cast<decltype(lib.builder_create)>(sym("ccu_builder_create"));
        lib.builder_destroy = reinterpret_cast<decltype(lib.builder_destroy)>(sym("ccu_builder_destroy"));
        lib.builder_const = reinterpret_cast<decltype(lib.builder_const)>(sym("ccu_builder_const"));
        lib.builder_input = reinterpret_cast<decltype(lib.builder_input)>(sym("ccu_builder_input"));
        lib.builder_op = reinterpret_cast<decltype(lib.builder_op)>(sym("ccu_builder_op"));
        lib.builder_output = reinterpret_cast<decltype(lib.builder_output)>(sym("ccu_builder_output"));
        lib.builder_ldl = reinterpret_cast<decltype(lib.builder_ldl)>(sym("ccu_builder_ldl"));
        lib.builder_qr = reinterpret_cast<decltype(lib.builder_qr)>(sym("ccu_builder_qr"));
        lib.builder_mtimes = reinterpret_cast<decltype(lib.builder_mtimes)>(sym("ccu_builder_mtimes"));
        lib.tape_create = reinterpret_cast<decltype(lib.tape_create)>(sym("ccu_tape_create"));
        lib.tape_destroy = reinterpret_cast<decltype(lib.tape_destroy)>(sym("ccu_tape_destroy"));
        lib.eval_reduce_device = reinterpret_cast<decltype(lib.eval_reduce_device)>(sym("ccu_map_eval_reduce_device"));
        lib.set_device = reinterpret_cast<decltype(lib.set_device)>(sym("ccu_set_device"));
        lib.dev_malloc = reinterpret_cast<decltype(lib.dev_malloc)>(sym("ccu_malloc"));
        lib.dev_free = reinterpret_cast<decltype(lib.dev_free)>(sym("ccu_free"));
        lib.memcpy_h2d = reinterpret_cast<decltype(lib.memcpy_h2d)>(sym("ccu_memcpy_h2d"));
        lib.memcpy_d2h = reinterpret_cast<decltype(lib.memcpy_d2h)>(sym("ccu_memcpy_d2h"));
        lib.stream_sync = reinterpret_cast<decltype(lib.stream_sync)>(sym("ccu_stream_sync"));
        lib.builder_select = reinterpret_cast<decltype(lib.builder_select)>(sym("ccu_builder_select"));
        lib.builder_export = reinterpret_cast<decltype(lib.builder_export)>(sym("ccu_builder_export"));
        if (!ok) {
          lib.error = "'" + name + "' does not export the casadi_cuda.h entry points";
          lib.handle = nullptr;
        } else if (lib.abi_version() != 3) {
          lib.error = "'" + name + "' has ABI version " + str(lib.abi_version()) + ", expected 3";
          lib.handle = nullptr;
        }
      });
      return lib;
    }

    // ------------------------------------------------------------------------------------------------
    // Lowering of an MX function to ONE scalar tape through the tape builder of libcasadi_cuda.so.
    // Walks the MX algorithm the way MXFunction::eval does (mx_function.cpp:435-490) and the way the
    // reference's own exporter reads it (mx_function.cpp:1620-1630: instruction_id/_MX/_input/_output),
    // keeping one value handle per nonzero of every work-vector element; each node is replayed with the
    // floating-point operations of its numeric eval, Linsol calls by tracing casadi_ldl / casadi_qr over
    // the (shared) pattern.  Only the vocabulary below is supported; anything else raises.
    // ------------------------------------------------------------------------------------------------
    typedef std::vector<ccu_int> Vals;

    struct Lowering {
      CudaLib& lib;
      void* b;
      std::map<unsigned long long, ccu_int> const_cache;
      ccu_int fail_count = -1;  // sum over QR solves of (nullity != 0)

      explicit Lowering(CudaLib& l) : lib(l), b(l.builder_create()) {}

      ccu_int cst(double v) {
        unsigned long long bits;
        std::memcpy(&bits, &v, 8);
        auto it = const_cache.find(bits);
        if (it != const_cache.end()) return it->second;
        ccu_int h = lib.builder_const(b, v);
        const_cache[bits] = h;
        return h;
      }
      ccu_int op(int o, ccu_int x, ccu_int y = -1) {
        ccu_int h = lib.builder_op(b, o, x, y);
        casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
        return h;
      }
      static std::vector<ccu_int> pattern(const Sparsity& sp) {
        std::vector<casadi_int> c = sp.compress();
        return std::vector<ccu_int>(c.begin(), c.end());
      }

      // Linsol::nfact + solve on handles: A = nonzeros of the matrix, xs = right-hand sides on entry, solutions on return
      void linsol_solve(const Linsol& ls, const Vals& A, Vals& xs, casadi_int nrhs, bool tr) {
        const Sparsity& sp = ls.sparsity();
        std::vector<ccu_int> spa = pattern(sp);
        if (ls.plugin_name() == "ldl") {
          // symbolic phase as in LinsolLdl::init (linsol_ldl.cpp:67-100, default options)
          std::vector<casadi_int> p;
          Sparsity lt = sp.ldl(p, true);
          std::vector<ccu_int> splt = pattern(lt), pp(p.begin(), p.end());
          casadi_assert(lib.builder_ldl(b, spa.data(), splt.data(), pp.data(), A.data(), xs.data(), nrhs, nullptr) == 0,
                        "Map 'cuda': " + std::string(lib.last_error()));
        } else if (ls.plugin_name() == "qr") {
          // symbolic phase as in LinsolQr::init (linsol_qr.cpp:67-84, default options, eps = 1e-12)
          Sparsity spv, spr;
          std::vector<casadi_int> prinv, pc;
          sp.qr_sparse(spv, spr, prinv, pc);
          std::vector<ccu_int> v1 = pattern(spv), r1 = pattern(spr), pi(prinv.begin(), prinv.end()),
                               pcc(pc.begin(), pc.end());
          ccu_int nullity = -1;
          casadi_assert(lib.builder_qr(b, spa.data(), v1.data(), r1.data(), pi.data(), pcc.data(), A.data(), xs.data(),
                                       nrhs, tr ? 1 : 0, 1e-12, &nullity) == 0,
                        "Map 'cuda': " + std::string(lib.last_error()));
          ccu_int bad = op(OP_NE, nullity, cst(0.));
          fail_count = fail_count < 0 ? bad : op(OP_ADD, fail_count, bad);
        } else if (ls.plugin_name() == "tridiag") {
          tridiag_solve(sp, A, xs, nrhs, tr);
        } else {
          casadi_error("Map 'cuda': linear solver plugin '" + ls.plugin_name()
                       + "' has no device implementation (supported: ldl, qr, tridiag)");
        }
      }

      // LinsolTridiag::solve (casadi/solvers/linsol_tridiag.cpp:85-151) on handles: the Thomas algorithm with the
      // plugin's own operation order -- the coefficients c (ctr for the transposed system) once per factorisation, then
      // d and the back substitution per right-hand side.  The plugin indexes A as A[colind[i] + offset] and so assumes a
      // tridiagonal pattern without checking it; here a different pattern is refused.  (The reference also computes
      // c[n-1] from one element past the end of A and never uses it: not traced.)
      void tridiag_solve(const Sparsity& sp, const Vals& A, Vals& xs, casadi_int nrhs, bool tr) {
        const casadi_int n = sp.size1();
        const casadi_int *ci = sp.colind(), *row = sp.row();
        bool ok = sp.size2() == n && n >= 2;
        for (casadi_int c = 0; ok && c < n; ++c) {
          const casadi_int lo = std::max<casadi_int>(c - 1, 0), hi = std::min<casadi_int>(c + 1, n - 1);
          ok = ci[c + 1] - ci[c] == hi - lo + 1;
          for (casadi_int k = ci[c]; ok && k < ci[c + 1]; ++k) ok = row[k] == lo + (k - ci[c]);
        }
        casadi_assert(ok, "Map 'cuda': linear solver 'tridiag' needs a square tridiagonal pattern with every band entry present, got "
                      + sp.dim(true));
        casadi_assert(static_cast<casadi_int>(xs.size()) == n * nrhs, "Map 'cuda': right-hand side size mismatch in 'tridiag'");
        auto a = [&](casadi_int k) { return A.at(k); };
        Vals c(n, -1), d(n, -1);
        // denominator of row i, as the plugin writes it in both loops
        auto denom = [&](casadi_int i) {
          if (tr) return op(OP_SUB, a(ci[i] + 1), op(OP_MUL, a(ci[i] + 0), c[i - 1]));
          return i == 1 ? op(OP_SUB, a(ci[1] + 1), op(OP_MUL, a(ci[0] + 1), c[0]))
                        : op(OP_SUB, a(ci[i] + 1), op(OP_MUL, a(ci[i - 1] + 2), c[i - 1]));
        };
        c[0] = tr ? op(OP_DIV, a(ci[0] + 1), a(ci[0] + 0)) : op(OP_DIV, a(ci[1] + 0), a(ci[0] + 0));
        for (casadi_int i = 1; i + 1 < n; ++i) {
          const ccu_int den = denom(i);
          c[i] = tr ? op(OP_DIV, a(ci[i] + 2), den) : op(OP_DIV, a(ci[i + 1] + 0), den);
        }
        for (casadi_int k = 0; k < nrhs; ++k) {
          ccu_int* x = xs.data() + k * n;
          d[0] = op(OP_DIV, x[0], a(ci[0] + 0));
          for (casadi_int i = 1; i < n; ++i) {
            const ccu_int den = denom(i);
            const ccu_int sub = tr ? a(ci[i] + 0) : (i == 1 ? a(ci[0] + 1) : a(ci[i - 1] + 2));
            d[i] = op(OP_DIV, op(OP_SUB, x[i], op(OP_MUL, sub, d[i - 1])), den);
          }
          x[n - 1] = d[n - 1];
          for (casadi_int i = n - 2; i >= 0; --i) x[i] = op(OP_SUB, d[i], op(OP_MUL, c[i], x[i + 1]));
        }
      }

      // Inline an SX function: replay its tape (sx_function.cpp:111-124) over handles
      void call_sx(const Function& f, const std::vector<const Vals*>& arg, std::vector<Vals*>& res) {
        std::vector<ccu_int> w(f.sz_w(), -1);
        casadi_int n = f.n_instructions();
        for (casadi_int k = 0; k < n; ++k) {
          casadi_int o = f.instruction_id(k);
          std::vector<casadi_int> in = f.instruction_input(k), out = f.instruction_output(k);
          if (o == OP_CONST) {
            w[out.at(0)] = cst(f.instruction_constant(k));
          } else if (o == OP_INPUT) {
            const Vals* a = arg.at(in.at(0));
            w[out.at(0)] = a ? a->at(in.at(1)) : cst(0.);
          } else if (o == OP_OUTPUT) {
            if (res.at(out.at(0))) res[out.at(0)]->at(out.at(1)) = w[in.at(0)];
          } else {
            casadi_assert(o != OP_CALL && o != OP_PARAMETER, "Map 'cuda': unsupported SX instruction in '" + f.name() + "'");
            w[out.at(0)] = op(static_cast<int>(o), w[in.at(0)], in.size() > 1 ? w[in.at(1)] : -1);
          }
        }
      }

      // Lower an MX function given the handles of its arguments; fills res (pre-sized by the caller)
      void call_mx(const Function& f, const std::vector<const Vals*>& arg, std::vector<Vals*>& res) {
        std::map<casadi_int, Vals> w;  // work-vector element -> handles
        casadi_int n = f.n_instructions();
        for (casadi_int k = 0; k < n; ++k) {
          casadi_int o = f.instruction_id(k);
          MX x = f.instruction_MX(k);
          std::vector<casadi_int> in = f.instruction_input(k), out = f.instruction_output(k);
          auto W = [&](casadi_int i) -> const Vals& {
            auto it = w.find(i);
            casadi_assert(it != w.end(), "Map 'cuda': MX work element read before it is written");
            return it->second;
          };
          if (o == OP_INPUT) {
            Dict inf = x.info();
            casadi_int ind = inf.at("ind"), off = inf.at("offset");
            Vals v(x.nnz());
            for (casadi_int e = 0; e < x.nnz(); ++e) v[e] = arg.at(ind) ? arg[ind]->at(off + e) : cst(0.);
            w[out.at(0)] = v;
          } else if (o == OP_OUTPUT) {
            Dict inf = x.info();
            casadi_int ind = inf.at("ind"), off = inf.at("offset");
            const Vals& v = W(in.at(0));
            if (res.at(ind)) for (size_t e = 0; e < v.size(); ++e) res[ind]->at(off + e) = v[e];
          } else if (o == OP_CONST) {
            DM v = static_cast<DM>(x);
            Vals r(v.nnz());
            for (casadi_int e = 0; e < v.nnz(); ++e) r[e] = cst(v.nonzeros()[e]);
            w[out.at(0)] = r;
          } else if (o == OP_MTIMES) {
            // Multiplication::eval_kernel = casadi_mtimes (multiplication.cpp:64-67); the dense variants call BLAS
            if (dynamic_cast<const DenseMultiplication*>(x.get()) != nullptr
                || dynamic_cast<const PseudoDenseMultiplication*>(x.get()) != nullptr
                || dynamic_cast<const DenseSparseMultiplication*>(x.get()) != nullptr) {
              // the dense variants run casadi_mtimes_dense / casadi_mtimes_dense_sparse for T = double and T = SXElem alike
              // (multiplication.cpp:172-258) as long as the node's BLAS backend is "reference" (Blas::mtimes, blas.cpp:201-210);
              // a BLAS plugin sums in its own order and has no device counterpart
              casadi_assert(blas_is_reference(x), "Map 'cuda': dense matrix products through a BLAS plugin ("
                            + x.class_name() + ") are not supported on the device (only the reference loops are)");
              node_sx(f, x, in, out, w);
              continue;
            }
            Vals z = W(in.at(0));
            const Vals &xx = W(in.at(1)), &yy = W(in.at(2));
            std::vector<ccu_int> spx = pattern(x.dep(1).sparsity()), spy = pattern(x.dep(2).sparsity()),
                                 spz = pattern(x.sparsity());
            casadi_assert(lib.builder_mtimes(b, xx.data(), spx.data(), yy.data(), spy.data(), z.data(), spz.data()) == 0,
                          "Map 'cuda': " + std::string(lib.last_error()));
            w[out.at(0)] = z;
          } else if (o == OP_SOLVE) {
            bool tr = x.info().at("tr");
            const Linsol* ls = nullptr;
            if (auto* n0 = dynamic_cast<const LinsolCall<false>*>(x.get())) ls = &n0->linsol_;
            if (auto* n1 = dynamic_cast<const LinsolCall<true>*>(x.get())) ls = &n1->linsol_;
            casadi_assert(ls != nullptr, "Map 'cuda': " + x.class_name() + " is not a Linsol call");
            Vals xs = W(in.at(0));  // right-hand sides, overwritten by the solutions (solve_impl.hpp:60)
            linsol_solve(*ls, W(in.at(1)), xs, x.dep(0).size2(), tr);
            w[out.at(0)] = xs;
          } else if (o == OP_CALL) {
            Function fc = x.which_function();
            std::vector<const Vals*> a(fc.n_in(), nullptr);
            for (casadi_int j = 0; j < fc.n_in(); ++j) if (in.at(j) >= 0) a[j] = &W(in[j]);
            std::vector<Vals> r(fc.n_out());
            std::vector<Vals*> rp(fc.n_out(), nullptr);
            for (casadi_int j = 0; j < fc.n_out(); ++j) {
              if (out.at(j) < 0) continue;
              r[j].assign(fc.nnz_out(j), cst(0.));
              rp[j] = &r[j];
            }
            call(fc, a, rp);
            for (casadi_int j = 0; j < fc.n_out(); ++j) if (out[j] >= 0) w[out[j]] = r[j];
          } else if (o == OP_HORZCAT || o == OP_VERTCAT || o == OP_DIAGCAT) {
            Vals r;  // Concat::eval_gen: the nonzeros of the arguments one after the other (concat.cpp)
            for (casadi_int i : in) { const Vals& v = W(i); r.insert(r.end(), v.begin(), v.end()); }
            w[out.at(0)] = r;
          } else if (o == OP_HORZSPLIT || o == OP_VERTSPLIT || o == OP_DIAGSPLIT) {
            std::vector<casadi_int> off = x.info().at("offset");  // Split::eval_gen (split.cpp)
            const Vals& v = W(in.at(0));
            casadi_int no = static_cast<casadi_int>(out.size());
            for (casadi_int j = 0; j < no; ++j) {
              if (out[j] < 0) continue;
              casadi_int nz_first = off.at(j), nz_last = off.at(j + 1);
              w[out[j]] = Vals(v.begin() + nz_first, v.begin() + nz_last);
            }
          } else if (o == OP_RESHAPE) {
            w[out.at(0)] = W(in.at(0));
          } else if (o == OP_GETNONZEROS) {
            Dict inf = x.info();
            if (inf.find("nz") == inf.end()) {  // the slice forms (GetNonzerosSlice / Slice2, getnonzeros.cpp)
              node_sx(f, x, in, out, w);
              continue;
            }
            std::vector<casadi_int> nz = inf.at("nz");
            const Vals& v = W(in.at(0));
            Vals r(nz.size());
            for (size_t e = 0; e < nz.size(); ++e) r[e] = nz[e] >= 0 ? v.at(nz[e]) : cst(0.);
            w[out.at(0)] = r;
          } else if (o == OP_HORZREPMAT) {
            const Vals& v = W(in.at(0));  // HorzRepmat::eval_gen (repmat.cpp:44-50)
            Vals r;
            casadi_int reps = v.empty() ? 0 : x.nnz() / static_cast<casadi_int>(v.size());
            for (casadi_int i = 0; i < reps; ++i) r.insert(r.end(), v.begin(), v.end());
            w[out.at(0)] = r;
          } else if (o == OP_HORZREPSUM) {
            const Vals& v = W(in.at(0));  // HorzRepsum::eval_gen (repmat.cpp:127-135): zero, then += in order
            casadi_int nnz = x.nnz(), reps = nnz ? static_cast<casadi_int>(v.size()) / nnz : 0;
            Vals r(nnz, cst(0.));
            for (casadi_int i = 0; i < reps; ++i)
              for (casadi_int e = 0; e < nnz; ++e) r[e] = op(OP_ADD, r[e], v[i * nnz + e]);
            w[out.at(0)] = r;
          } else if (x.n_dep() == 2 && x.is_binary()) {
            // BinaryMX<ScX,ScY>::eval_gen (binary_mx.cpp): element-wise, scalars broadcast
            const Vals &a = W(in.at(0)), &c = W(in.at(1));
            casadi_int nn = x.nnz();
            bool sa = a.size() == 1 && nn != 1, sc = c.size() == 1 && nn != 1;
            casadi_assert((sa || static_cast<casadi_int>(a.size()) == nn) && (sc || static_cast<casadi_int>(c.size()) == nn),
                          "Map 'cuda': operand patterns of " + x.class_name() + " do not match its result");
            Vals r(nn);
            for (casadi_int e = 0; e < nn; ++e) r[e] = op(static_cast<int>(o), a[sa ? 0 : e], c[sc ? 0 : e]);
            w[out.at(0)] = r;
          } else if (x.n_dep() == 1 && x.is_unary()) {
            const Vals& a = W(in.at(0));
            Vals r(a.size());
            for (size_t e = 0; e < a.size(); ++e) r[e] = op(static_cast<int>(o), a[e]);
            w[out.at(0)] = r;
          } else if (o == OP_BSPLINE) {
            const BSplineCommon* bs = dynamic_cast<const BSplineCommon*>(x.get());
            casadi_assert(bs != nullptr, "Map 'cuda': " + x.class_name() + " is not a B-spline node");
            const BSpline* bc = dynamic_cast<const BSpline*>(x.get());
            Vals coeffs;
            if (bc) { for (double v : bc->coeffs_) coeffs.push_back(cst(v)); } else { coeffs = W(in.at(1)); }
            w[out.at(0)] = bspline_eval(f, bs, W(in.at(0)), coeffs);
          } else if (o == OP_DOT || o == OP_NORMF || o == OP_NORM1 || o == OP_NORMINF) {
            // casadi_dot / casadi_norm_2 / casadi_norm_1 / casadi_norm_inf (runtime/): the accumulator starts from an
            // explicit zero, as in the numeric evaluation (0 + (-0) is +0; an SX expansion would drop the addition)
            const Vals& a = W(in.at(0));
            const Vals& c = o == OP_DOT ? W(in.at(1)) : a;
            casadi_assert(a.size() == c.size(), "Map 'cuda': operand sizes of " + x.class_name() + " differ");
            ccu_int r = cst(0.);
            for (size_t e = 0; e < a.size(); ++e) {
              if (o == OP_NORM1) r = op(OP_ADD, r, op(OP_FABS, a[e]));
              else if (o == OP_NORMINF) r = op(OP_FMAX, r, op(OP_FABS, a[e]));
              else r = op(OP_ADD, r, op(OP_MUL, a[e], c[e]));
            }
            if (o == OP_NORMF) r = op(OP_SQRT, r);
            w[out.at(0)] = Vals(1, r);
          } else if (o == OP_ASSERTION) {
            // Assertion::eval (assertion.cpp:69-79): the value passes through; the evaluation fails unless the condition is
            // exactly 1.  An instance that violates it is counted with the failed QR factorisations: the map then fails like
            // the reference's (whose serial map raises "Assertion error"), instead of dropping the check as an expansion does
            const ccu_int bad = op(OP_NE, W(in.at(1)).at(0), cst(1.));
            fail_count = fail_count < 0 ? bad : op(OP_ADD, fail_count, bad);
            w[out.at(0)] = W(in.at(0));
          } else if (o == OP_LOGSUMEXP) {
            // casadi_logsumexp (runtime/casadi_logsumexp.hpp): "max" is the last x[i] that exceeds x[0] (sic: the reference
            // compares with x[0], not with the running maximum), the sum skips that element, log1p(sum) + max.  The arg-max
            // is data: one-hot flags and selects.  (LogSumExp has no eval_sx; its numeric evaluation is replayed here.)
            const Vals& a = W(in.at(0));
            const casadi_int nn = static_cast<casadi_int>(a.size());
            casadi_assert(nn >= 1, "Map 'cuda': logsumexp of an empty operand");
            ccu_int r = a[0];
            if (nn > 1) {
              const ccu_int zero = cst(0.);
              auto sel = [&](ccu_int c, ccu_int p, ccu_int q) {
                ccu_int h = lib.builder_select(b, c, p, q);
                casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
                return h;
              };
              Vals gt(nn, zero), is_max(nn, zero);
              ccu_int mx = a[0];
              for (casadi_int i = 1; i < nn; ++i) { gt[i] = op(OP_LT, a[0], a[i]); mx = sel(gt[i], a[i], mx); }
              ccu_int later = zero;  // some j > i exceeds x[0]
              for (casadi_int i = nn; i-- > 1; ) { is_max[i] = op(OP_AND, gt[i], op(OP_NOT, later)); later = op(OP_OR, later, gt[i]); }
              is_max[0] = op(OP_NOT, later);
              ccu_int sum = zero;
              for (casadi_int i = 0; i < nn; ++i) sum = sel(is_max[i], sum, op(OP_ADD, sum, op(OP_EXP, op(OP_SUB, a[i], mx))));
              r = op(OP_ADD, op(OP_LOG1P, sum), mx);
            }
            w[out.at(0)] = Vals(1, r);
          } else if (o == OP_BILIN) {
            // casadi_bilin (runtime/casadi_bilin.hpp): ret = 0; ret += x[rr]*A[el]*y[cc] column by column
            const Vals &A = W(in.at(0)), &xx = W(in.at(1)), &yy = W(in.at(2));
            const Sparsity spA = x.dep(0).sparsity();
            const casadi_int *colind = spA.colind(), *row = spA.row();
            ccu_int r = cst(0.);
            for (casadi_int cc = 0; cc < spA.size2(); ++cc)
              for (casadi_int el = colind[cc]; el < colind[cc + 1]; ++el)
                r = op(OP_ADD, r, op(OP_MUL, op(OP_MUL, xx.at(row[el]), A.at(el)), yy.at(cc)));
            w[out.at(0)] = Vals(1, r);
          } else if (o == OP_TRANSPOSE || o == OP_PROJECT || o == OP_RANK1
                     || o == OP_SETNONZEROS || o == OP_ADDNONZEROS || o == OP_MMIN || o == OP_MMAX || o == OP_SPARSITY_CAST
                     || o == OP_LIFT || o == OP_EINSTEIN) {
            node_sx(f, x, in, out, w);
          } else {
            casadi_error("Map 'cuda': MX operation '" + x.class_name() + "' (op " + str(o) + ") in function '" + f.name()
                         + "' has no device lowering");
          }
        }
      }

      // true when a Multiplication node multiplies with the reference loops (blas_shorthand_ == 0, multiplication.hpp:207)
      static bool blas_is_reference(const MX& x) {
        struct Peek : public Multiplication { using Multiplication::blas_shorthand_; };
        casadi_int Multiplication::* member = &Peek::blas_shorthand_;
        const Multiplication* m = dynamic_cast<const Multiplication*>(x.get());
        return m != nullptr && m->*member == 0;
      }

      // A node whose numeric eval and eval_sx instantiate ONE template (eval_gen<T>, casadi_*<T1> of the runtime): the
      // reference's own symbolic evaluation of the node on fresh symbols -- exactly what Function::expand() runs for it
      // (MXFunction::eval_sx, mx_function.cpp:1229-1278) -- wrapped into an SX function and inlined like any other.
      // Pure data movement (transpose, project, get/set nonzeros) records nothing.  As in an expansion, SX drops the
      // additions of the constant zeros a node starts its accumulators from (0 + a*b is a*b): the sign of a zero result
      // may differ from the numeric eval, nothing else.
      void node_sx(const Function& f, const MX& x, const std::vector<casadi_int>& in, const std::vector<casadi_int>& out,
                   std::map<casadi_int, Vals>& w) {
        const casadi_int nd = x.n_dep(), no = static_cast<casadi_int>(out.size());
        std::vector<SX> sa(nd), so(no);
        std::vector<const SXElem*> argp(std::max<size_t>(x->sz_arg(), nd) + 1, nullptr);
        std::vector<SXElem*> resp(std::max<size_t>(x->sz_res(), no) + 1, nullptr);
        std::vector<SX> fin, fout;
        std::vector<const Vals*> a;
        for (casadi_int d = 0; d < nd; ++d) {
          if (in.at(d) < 0) continue;
          sa[d] = SX::sym("a" + str(d), x.dep(d).sparsity());
          argp[d] = sa[d].ptr();
          fin.push_back(sa[d]);
          auto it = w.find(in[d]);
          casadi_assert(it != w.end(), "Map 'cuda': MX work element read before it is written");
          casadi_assert(static_cast<casadi_int>(it->second.size()) == x.dep(d).nnz(), "Map 'cuda': operand of " + x.class_name()
                        + " has " + str(it->second.size()) + " nonzeros, expected " + str(x.dep(d).nnz()));
          a.push_back(&it->second);
        }
        for (casadi_int k = 0; k < no; ++k) {
          if (out[k] < 0) continue;
          so[k] = SX::zeros(x->sparsity(k));
          resp[k] = so[k].ptr();
        }
        std::vector<casadi_int> iw(x->sz_iw() + 1);
        std::vector<SXElem> ww(x->sz_w() + 1);
        int flag = 1;
        std::string why;
        try {
          flag = x->eval_sx(argp.data(), resp.data(), iw.data(), ww.data());
        } catch (std::exception& e) {
          why = e.what();
        }
        casadi_assert(flag == 0, "Map 'cuda': MX operation '" + x.class_name() + "' in function '" + f.name()
                      + "' has no device lowering (its symbolic evaluation failed" + (why.empty() ? "" : ": " + why) + ")");
        std::vector<Vals> r;
        std::vector<casadi_int> which;
        for (casadi_int k = 0; k < no; ++k) {
          if (out[k] < 0) continue;
          fout.push_back(so[k]);
          which.push_back(k);
        }
        r.resize(which.size());
        std::vector<Vals*> rp(which.size());
        for (size_t q = 0; q < which.size(); ++q) {
          r[q].assign(fout[q].nnz(), cst(0.));
          rp[q] = &r[q];
        }
        Function g("node_" + str(x.op()), fin, fout);
        call_sx(g, a, rp);
        for (size_t q = 0; q < which.size(); ++q) w[out[which[q]]] = r[q];
      }

      // A fixed-step integrator with an explicit step function (the "rk" plugin, runge_kutta.cpp:68-135): replays
      // Integrator::eval (integrator.cpp:354-530) for the case without events and without backward states.  With the
      // time grid fixed, every step time t_k + j*h and step length h is a host constant computed by the reference's own
      // expressions (FixedStepIntegrator::advance_noevent, integrator.cpp:2034-2073), and the evaluation is nt * nj
      // calls of the discrete-time function "step" (and of its forward-mode function for an augmented integrator,
      // FixedStepIntegrator::stepF, :2119-2150) with the quadrature accumulated as q = qf + 1.*q_prev: a straight-line
      // program whose loop structure the tape re-rolling pass of libcasadi_cuda.so recovers.
      // Controls: interval k uses u[k] -- the reference keeps the previous interval's control while it compares equal
      // (Integrator::next_stop, :2566-2583), which differs only when consecutive controls are zeros of opposite sign.
      void call_fixed_step(const Function& f, const FixedStepIntegrator* I, const std::vector<const Vals*>& arg,
                           std::vector<Vals*>& res) {
        const std::string who = "Map 'cuda': integrator '" + f.name() + "' (" + f.class_name() + "): ";
        casadi_assert(I->has_function("step"), who + "implicit step functions (collocation) have no device lowering");
        casadi_assert(I->ne_ == 0, who + "events (zero-crossing functions) have no device lowering");
        casadi_assert(I->nz_ == 0 && I->nrz_ == 0, who + "algebraic variables have no device lowering");
        const Function& F = I->get_function("step");
        const casadi_int nfwd = I->nfwd_;
        Function dF;
        if (nfwd > 0) dF = I->get_function(FunctionInternal::forward_name("step", nfwd));
        const casadi_int nx = I->nx_, nq = I->nq_, np = I->np_, nu = I->nu_, nv = F.nnz_out(STEP_VF) * (1 + nfwd);
        const casadi_int nx1 = I->nx1_, nq1 = I->nq1_, np1 = I->np1_, nu1 = I->nu1_, nv1 = F.nnz_out(STEP_VF);
        const casadi_int nt = I->nt();
        const ccu_int zero = cst(0.);
        auto take = [&](casadi_int j, casadi_int off, casadi_int n) {  // casadi_copy with a null source clears
          Vals r(n, zero);
          if (arg.at(j)) for (casadi_int i = 0; i < n; ++i) r[i] = arg[j]->at(off + i);
          return r;
        };
        Vals x = take(INTEGRATOR_X0, 0, nx), p = take(INTEGRATOR_P, 0, np), q(nq, zero);
        Vals v(nv, cst(std::numeric_limits<double>::quiet_NaN()));  // FixedStepIntegrator::reset, :2230
        // the state after every step, for the backward sweep (m->x_tape, m->v_tape; advance_noevent :2066-2070)
        const bool backward = I->nrx_ > 0;
        std::vector<Vals> x_tape, v_tape;
        if (backward) x_tape.push_back(x);
        double t = I->t0_;
        Vals u(nu, zero), u_raw_prev;
        for (casadi_int k = 0; k < nt; ++k) {
          const double t_next = I->tout_.at(k);
          // Integrator::eval passes new controls only where next_stop (:2566-2583) finds u[k-1] != u[k] in some entry, and
          // otherwise keeps what it holds -- equal values, but possibly a zero of the other sign: a bit-exact select
          const Vals u_raw = take(INTEGRATOR_U, k * nu, nu);
          if (k == 0 || nu == 0 || !arg.at(INTEGRATOR_U)) {
            u = u_raw;
          } else {
            ccu_int changed = zero;
            for (casadi_int i = 0; i < nu; ++i) changed = op(OP_OR, changed, op(OP_NE, u_raw_prev[i], u_raw[i]));
            for (casadi_int i = 0; i < nu; ++i) {
              ccu_int h = lib.builder_select(b, changed, u_raw[i], u[i]);
              casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
              u[i] = h;
            }
          }
          u_raw_prev = u_raw;
          const casadi_int nj = I->disc_.at(k + 1) - I->disc_.at(k);
          const double h = (t_next - t) / nj;
          for (casadi_int j = 0; j < nj; ++j) {
            const double tj = t + j * h;
            const Vals x_prev = x, v_prev = v, q_prev = q;
            const Vals tv(1, cst(tj)), hv(1, cst(h));
            const Vals x0(x_prev.begin(), x_prev.begin() + nx1), v0(v_prev.begin(), v_prev.begin() + nv1),
                       pp(p.begin(), p.begin() + np1), uu(u.begin(), u.begin() + nu1);
            Vals xf(nx1, zero), vf(nv1, zero), qf(nq1, zero);
            {
              std::vector<const Vals*> a(F.n_in(), nullptr);
              a[STEP_T] = &tv; a[STEP_H] = &hv; a[STEP_X0] = &x0; a[STEP_V0] = &v0; a[STEP_P] = &pp; a[STEP_U] = &uu;
              std::vector<Vals*> r(F.n_out(), nullptr);
              r[STEP_XF] = &xf; r[STEP_VF] = &vf; r[STEP_QF] = &qf;
              call(F, a, r);
            }
            std::copy(xf.begin(), xf.end(), x.begin());
            std::copy(vf.begin(), vf.end(), v.begin());
            std::copy(qf.begin(), qf.end(), q.begin());
            if (nfwd > 0) {
              const Vals fx0(x_prev.begin() + nx1, x_prev.end()), fv0(v_prev.begin() + nv1, v_prev.end()),
                         fp(p.begin() + np1, p.end()), fu(u.begin() + nu1, u.end());
              Vals fxf(nx - nx1, zero), fvf(nv - nv1, zero), fqf(nq - nq1, zero);
              std::vector<const Vals*> a(dF.n_in(), nullptr);
              a[STEP_T] = &tv; a[STEP_H] = &hv; a[STEP_X0] = &x0; a[STEP_V0] = &v0; a[STEP_P] = &pp; a[STEP_U] = &uu;
              a[STEP_NUM_IN + STEP_XF] = &xf; a[STEP_NUM_IN + STEP_VF] = &vf; a[STEP_NUM_IN + STEP_QF] = &qf;
              a[STEP_NUM_IN + STEP_NUM_OUT + STEP_X0] = &fx0; a[STEP_NUM_IN + STEP_NUM_OUT + STEP_V0] = &fv0;
              a[STEP_NUM_IN + STEP_NUM_OUT + STEP_P] = &fp; a[STEP_NUM_IN + STEP_NUM_OUT + STEP_U] = &fu;
              std::vector<Vals*> r(dF.n_out(), nullptr);
              r[STEP_XF] = &fxf; r[STEP_VF] = &fvf; r[STEP_QF] = &fqf;
              call(dF, a, r);
              std::copy(fxf.begin(), fxf.end(), x.begin() + nx1);
              std::copy(fvf.begin(), fvf.end(), v.begin() + nv1);
              std::copy(fqf.begin(), fqf.end(), q.begin() + nq1);
            }
            // casadi_axpy(nq_, 1., q_prev, q)
            const ccu_int one = cst(1.);
            for (casadi_int i = 0; i < nq; ++i) q[i] = op(OP_ADD, q[i], op(OP_MUL, one, q_prev[i]));
            if (backward) { x_tape.push_back(x); v_tape.push_back(v); }
          }
          t = t_next;
          if (res.at(INTEGRATOR_XF)) std::copy(x.begin(), x.end(), res[INTEGRATOR_XF]->begin() + k * nx);
          if (res.at(INTEGRATOR_QF)) std::copy(q.begin(), q.end(), res[INTEGRATOR_QF]->begin() + k * nq);
        }
        if (backward) backward_sweep(f, I, arg, res, x_tape, v_tape, p);
      }

      // The backward integration of Integrator::eval (integrator.cpp:459-517) for a fixed-step integrator with backward
      // states (the adjoint integrator Integrator::get_reverse creates): resetB, then from the last output time to t0 the
      // impulse of the adjoint seeds (impulseB, :2254-2271) and the backward steps of the interval (retreat, :2081-2117;
      // stepB :2152-2176 calls the plugin's adj<nadj>_step with the states the forward sweep left on its tape).  The
      // reference adds an impulse only when a seed of that output time is nonzero and integrates backward only once some
      // impulse has occurred; per instance of a map these are data, so both become bit-exact selects on the 0/1 flags.
      void backward_sweep(const Function& f, const FixedStepIntegrator* I, const std::vector<const Vals*>& arg,
                          std::vector<Vals*>& res, const std::vector<Vals>& x_tape, const std::vector<Vals>& v_tape,
                          const Vals& p) {
        const casadi_int nadj = I->nadj_, nfwd = I->nfwd_, nrx = I->nrx_, nrq = I->nrq_, nuq = I->nuq_, nrp = I->nrp_, nrv = I->nrv_;
        const casadi_int nu = I->nu_, nt = I->nt();
        // sizes of the nominal block of every vector (the forward sensitivities of an augmented integrator follow it)
        const casadi_int nx1 = I->nx1_, np1 = I->np1_, nu1 = I->nu1_, nv1 = I->nv1_, nrv1 = I->nrv1_;
        const casadi_int nrx1 = I->nrx1_ * nadj, nrq1 = I->nrq1_ * nadj, nuq1 = I->nuq1_ * nadj, nrp1 = I->nrp1_ * nadj;
        const Function& B = I->get_function(FunctionInternal::reverse_name("step", nadj));
        Function dB;
        if (nfwd > 0) dB = I->get_function(FunctionInternal::forward_name(FunctionInternal::reverse_name("step", nadj), nfwd));
        auto head = [](const Vals& v, casadi_int n) { return Vals(v.begin(), v.begin() + n); };
        auto tail = [](const Vals& v, casadi_int n) { return Vals(v.begin() + n, v.end()); };
        const ccu_int zero = cst(0.), one = cst(1.), minus_one = cst(-1.);
        auto take = [&](casadi_int j, casadi_int off, casadi_int n) {
          Vals r(n, zero);
          if (arg.at(j)) for (casadi_int i = 0; i < n; ++i) r[i] = arg[j]->at(off + i);
          return r;
        };
        auto sel = [&](ccu_int c, ccu_int a, ccu_int b2) {
          ccu_int h = lib.builder_select(b, c, a, b2);
          casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
          return h;
        };
        // resetB (:2239-2252)
        Vals adj_q(nrp, zero), adj_x(nrx, zero), adj_p(nrq, zero), adj_u(nuq, zero);
        const Vals rv(nrv, zero);  // (only algebraic variables put an impulse on it, and it is cleared after every step)
        ccu_int any_impulse = zero;
        std::vector<Vals> adj_u_out(nt, Vals(nuq, zero));
        for (casadi_int k = nt; k-- > 0; ) {
          const double t = I->tout_.at(k), t_next = k == 0 ? I->t0_ : I->tout_.at(k - 1);
          const Vals seed_x = take(INTEGRATOR_ADJ_XF, k * nrx, nrx), seed_q = take(INTEGRATOR_ADJ_QF, k * nrp, nrp);
          const Vals u = take(INTEGRATOR_U, k * nu, nu);
          // !all_zero(adj_xf) || !all_zero(rp)   (:2757-2766: v[i] != 0.)
          ccu_int nz = zero;
          if (arg.at(INTEGRATOR_ADJ_XF)) for (casadi_int i = 0; i < nrx; ++i) nz = op(OP_OR, nz, op(OP_NE, seed_x[i], zero));
          if (arg.at(INTEGRATOR_ADJ_QF)) for (casadi_int i = 0; i < nrp; ++i) nz = op(OP_OR, nz, op(OP_NE, seed_q[i], zero));
          // impulseB: casadi_axpy(n, 1., seed, state), only when a seed is nonzero
          if (arg.at(INTEGRATOR_ADJ_QF))
            for (casadi_int i = 0; i < nrp; ++i) adj_q[i] = sel(nz, op(OP_ADD, adj_q[i], op(OP_MUL, one, seed_q[i])), adj_q[i]);
          if (arg.at(INTEGRATOR_ADJ_XF))
            for (casadi_int i = 0; i < nrx; ++i) adj_x[i] = sel(nz, op(OP_ADD, adj_x[i], op(OP_MUL, one, seed_x[i])), adj_x[i]);
          any_impulse = op(OP_OR, any_impulse, nz);
          // retreat, as if an impulse had occurred; what it leaves is selected below
          Vals rx = adj_x, rp = adj_p, ru = adj_u;
          const casadi_int nj = I->disc_.at(k + 1) - I->disc_.at(k);
          const double h = (t - t_next) / nj;
          for (casadi_int j = nj; j-- > 0; ) {
            const double tj = t_next + j * h;
            const Vals x_prev = rx, p_prev = rp, u_prev = ru;
            const casadi_int tapeind = I->disc_.at(k) + j;
            const Vals tv(1, cst(tj)), hv(1, cst(h));
            // nominal blocks of the operands (stepB :2155-2176)
            const Vals x0 = head(x_tape.at(tapeind), nx1), xf = head(x_tape.at(tapeind + 1), nx1), vf = head(v_tape.at(tapeind), nv1),
                       pp = head(p, np1), uu = head(u, nu1), sx = head(x_prev, nrx1), srv = head(rv, nrv1), sq = head(adj_q, nrp1);
            std::vector<const Vals*> a(B.n_in(), nullptr);
            a[BSTEP_T] = &tv; a[BSTEP_H] = &hv; a[BSTEP_X0] = &x0; a[BSTEP_P] = &pp; a[BSTEP_U] = &uu;
            a[BSTEP_OUT_XF] = &xf; a[BSTEP_OUT_VF] = &vf;
            a[BSTEP_ADJ_XF] = &sx; a[BSTEP_ADJ_VF] = &srv; a[BSTEP_ADJ_QF] = &sq;
            // (outputs that are structurally empty in adj_step stay zero: issue #3353, :2171-2174)
            Vals ox(nrx1, zero), opar(nrq1, zero), ou(nuq1, zero);
            std::vector<Vals*> r(B.n_out(), nullptr);
            r[BSTEP_ADJ_X0] = &ox; r[BSTEP_ADJ_P] = &opar; r[BSTEP_ADJ_U] = &ou;
            call(B, a, r);
            if (nfwd > 0) {
              // forward sensitivities of the backward step (:2178-2217): the nominal operands and results, then the seeds
              const Vals fx0 = tail(x_tape.at(tapeind), nx1), fxf = tail(x_tape.at(tapeind + 1), nx1), fvf = tail(v_tape.at(tapeind), nv1),
                         fp = tail(p, np1), fu = tail(u, nu1), fsx = tail(x_prev, nrx1), fsrv = tail(rv, nrv1), fsq = tail(adj_q, nrp1);
              std::vector<const Vals*> fa(dB.n_in(), nullptr);
              for (casadi_int q = 0; q < BSTEP_NUM_IN; ++q) fa[q] = a[q];
              fa[BSTEP_NUM_IN + BSTEP_ADJ_X0] = &ox; fa[BSTEP_NUM_IN + BSTEP_ADJ_P] = &opar; fa[BSTEP_NUM_IN + BSTEP_ADJ_U] = &ou;
              const casadi_int o = BSTEP_NUM_IN + BSTEP_NUM_OUT;
              fa[o + BSTEP_X0] = &fx0; fa[o + BSTEP_P] = &fp; fa[o + BSTEP_U] = &fu; fa[o + BSTEP_OUT_XF] = &fxf; fa[o + BSTEP_OUT_VF] = &fvf;
              fa[o + BSTEP_ADJ_XF] = &fsx; fa[o + BSTEP_ADJ_VF] = &fsrv; fa[o + BSTEP_ADJ_QF] = &fsq;
              Vals fox(nrx - nrx1, zero), fopar(nrq - nrq1, zero), fou(nuq - nuq1, zero);
              std::vector<Vals*> fr(dB.n_out(), nullptr);
              fr[BSTEP_ADJ_X0] = &fox; fr[BSTEP_ADJ_P] = &fopar; fr[BSTEP_ADJ_U] = &fou;
              call(dB, fa, fr);
              ox.insert(ox.end(), fox.begin(), fox.end());
              opar.insert(opar.end(), fopar.begin(), fopar.end());
              ou.insert(ou.end(), fou.begin(), fou.end());
            }
            rx = ox;
            for (casadi_int i = 0; i < nrq; ++i) rp[i] = op(OP_ADD, opar[i], op(OP_MUL, one, p_prev[i]));
            for (casadi_int i = 0; i < nuq; ++i) ru[i] = op(OP_ADD, ou[i], op(OP_MUL, one, u_prev[i]));
          }
          // if (any_impulse) the states are what retreat left and adj_u of this interval is the running sum; otherwise the
          // states are untouched and the output is cleared
          for (casadi_int i = 0; i < nrx; ++i) adj_x[i] = sel(any_impulse, rx[i], adj_x[i]);
          for (casadi_int i = 0; i < nrq; ++i) adj_p[i] = sel(any_impulse, rp[i], adj_p[i]);
          for (casadi_int i = 0; i < nuq; ++i) adj_u[i] = sel(any_impulse, ru[i], adj_u[i]);
          for (casadi_int i = 0; i < nuq; ++i) adj_u_out[k][i] = sel(any_impulse, adj_u[i], zero);
          if (k == 0) {
            if (res.at(INTEGRATOR_ADJ_X0)) for (casadi_int i = 0; i < nrx; ++i) res[INTEGRATOR_ADJ_X0]->at(i) = sel(any_impulse, adj_x[i], zero);
            if (res.at(INTEGRATOR_ADJ_P)) for (casadi_int i = 0; i < nrq; ++i) res[INTEGRATOR_ADJ_P]->at(i) = sel(any_impulse, adj_p[i], zero);
          }
        }
        // adj_u per grid point, not cumulative (:2509-2515): adj_u[k] += -1 * adj_u[k+1], front to back
        for (casadi_int k = 0; k + 1 < nt; ++k)
          for (casadi_int i = 0; i < nuq; ++i)
            adj_u_out[k][i] = op(OP_ADD, adj_u_out[k][i], op(OP_MUL, minus_one, adj_u_out[k + 1][i]));
        if (res.at(INTEGRATOR_ADJ_U))
          for (casadi_int k = 0; k < nt; ++k) std::copy(adj_u_out[k].begin(), adj_u_out[k].end(), res[INTEGRATOR_ADJ_U]->begin() + k * nuq);
        (void)f;
      }

      // n evaluations of g over consecutive blocks of the operands (Map / MapSum); reduce_in: one shared block,
      // reduce_out: zero, then += the instance's result in index order
      void call_repeated(const Function& g, casadi_int n, const std::vector<bool>& reduce_in, const std::vector<bool>& reduce_out,
                         const std::vector<const Vals*>& arg, std::vector<Vals*>& res) {
        const casadi_int n_in = g.n_in(), n_out = g.n_out();
        for (casadi_int j = 0; j < n_out; ++j)
          if (res.at(j) && reduce_out.at(j)) res[j]->assign(g.nnz_out(j), cst(0.));
        for (casadi_int i = 0; i < n; ++i) {
          std::vector<Vals> a(n_in), r(n_out);
          std::vector<const Vals*> ap(n_in, nullptr);
          std::vector<Vals*> rp(n_out, nullptr);
          for (casadi_int j = 0; j < n_in; ++j) {
            if (!arg.at(j)) continue;
            const casadi_int nz = g.nnz_in(j), off = reduce_in.at(j) ? 0 : i * nz;
            casadi_assert(static_cast<casadi_int>(arg[j]->size()) >= off + nz, "Map 'cuda': operand " + str(j) + " of the embedded map '"
                          + g.name() + "' is too short");
            a[j].assign(arg[j]->begin() + off, arg[j]->begin() + off + nz);
            ap[j] = &a[j];
          }
          for (casadi_int j = 0; j < n_out; ++j) {
            if (!res.at(j)) continue;
            r[j].assign(g.nnz_out(j), cst(0.));
            rp[j] = &r[j];
          }
          call(g, ap, rp);
          for (casadi_int j = 0; j < n_out; ++j) {
            if (!res[j]) continue;
            const casadi_int nz = g.nnz_out(j);
            if (reduce_out.at(j)) {
              for (casadi_int e = 0; e < nz; ++e) res[j]->at(e) = op(OP_ADD, res[j]->at(e), r[j][e]);
            } else {
              for (casadi_int e = 0; e < nz; ++e) res[j]->at(i * nz + e) = r[j][e];
            }
          }
        }
      }

      // A B-spline node (MX::bspline, interpolant(..., "bspline", ...); BSpline::eval / BSplineParametric::eval, bspline.cpp:437-455):
      // casadi_nd_boor_eval (runtime/) replayed -- per dimension the knot span L (casadi_low), start = min(L, n_b-degree-1),
      // the initial basis vector from the comparisons of x with the knots, casadi_de_boor over the 2*degree+2 knots from
      // `start`, then casadi_tensor_ttv over the coefficients from `starts`.  Everything indexed by L or start is gathered
      // by bit-exact selects on one-hot flags; the `if (bottom)` guards of de Boor's recursion (repeated knots) are
      // selects on bottom != 0.
      Vals bspline_eval(const Function& f, const BSplineCommon* bs, const Vals& xin, const Vals& coeffs) {
        const std::string who = "Map 'cuda': B-spline in '" + f.name() + "': ";
        const casadi_int nd = static_cast<casadi_int>(bs->degree_.size()), m = bs->m_;
        casadi_assert(static_cast<casadi_int>(xin.size()) == nd, who + "unexpected argument size (batched evaluation has no device lowering)");
        const ccu_int zero = cst(0.), one = cst(1.);
        auto sel = [&](ccu_int c, ccu_int a, ccu_int b2) {
          ccu_int h = lib.builder_select(b, c, a, b2);
          casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
          return h;
        };
        double work = static_cast<double>(m);
        std::vector<std::vector<ccu_int>> start_hit(nd);  // one-hot flags of `start` per dimension
        std::vector<Vals> boor(nd);                        // the degree+1 basis values per dimension
        for (casadi_int k = 0; k < nd; ++k) {
          const casadi_int degree = bs->degree_[k], n_knots = bs->offset_[k + 1] - bs->offset_[k], n_b = n_knots - degree - 1;
          const double* knots = bs->knots_.data() + bs->offset_[k];
          const ccu_int xk = xin[k];
          // L = casadi_low(x, knots + degree, n_knots - 2*degree, mode): one-hot over 0 .. ng-2
          const double* g = knots + degree;
          const casadi_int ng = n_knots - 2 * degree;
          casadi_assert(ng >= 2, who + "too few knots");
          std::vector<ccu_int> L_hit(ng - 1, zero);
          if (bs->lookup_mode_.at(k) == 1) {
            const ccu_int t = op(OP_DIV, op(OP_MUL, op(OP_SUB, xk, cst(g[0])), cst(static_cast<double>(ng - 1))), cst(g[ng - 1] - g[0]));
            std::vector<ccu_int> ge(ng, zero);
            for (casadi_int j = 1; j <= ng - 2; ++j) ge[j] = op(OP_LE, cst(static_cast<double>(j)), t);
            for (casadi_int j = 0; j <= ng - 2; ++j)
              L_hit[j] = op(OP_AND, j == 0 ? one : ge[j], op(OP_NOT, j == ng - 2 ? zero : ge[j + 1]));
          } else if (bs->lookup_mode_.at(k) == 2) {
            // binary search (casadi_low case 2): x < g[1] -> 0, x > g[ng-1] -> ng-2, else the bisection -- on a
            // non-decreasing grid with possibly repeated points it ends at the LAST j with g[j] <= x
            std::vector<ccu_int> lt(ng, zero);
            for (casadi_int j = 1; j <= ng - 1; ++j) lt[j] = op(OP_LT, xk, cst(g[j]));
            for (casadi_int j = 0; j <= ng - 2; ++j)
              L_hit[j] = op(OP_AND, j == ng - 2 ? one : lt[j + 1], j == 0 ? one : op(OP_NOT, lt[j]));
          } else {
            std::vector<ccu_int> lt(ng, zero);
            for (casadi_int j = 1; j <= ng - 2; ++j) lt[j] = op(OP_LT, xk, cst(g[j]));
            for (casadi_int j = 0; j <= ng - 2; ++j)
              L_hit[j] = op(OP_AND, j == ng - 2 ? one : lt[j + 1], j == 0 ? one : op(OP_NOT, lt[j]));
          }
          // start = min(L, n_b - degree - 1): the flags of L folded onto 0 .. smax
          const casadi_int smax = n_b - degree - 1;
          casadi_assert(smax >= 0, who + "too few knots for the degree");
          start_hit[k].assign(smax + 1, zero);
          for (casadi_int j = 0; j <= ng - 2; ++j) {
            const casadi_int sidx = std::min(j, smax);
            start_hit[k][sidx] = start_hit[k][sidx] == zero ? L_hit[j] : op(OP_OR, start_hit[k][sidx], L_hit[j]);
          }
          auto gather_L = [&](casadi_int shift) {  // knots[L + shift]
            ccu_int v = cst(knots[(ng - 2) + shift]);
            for (casadi_int j = ng - 2; j-- > 0; ) v = sel(L_hit[j], cst(knots[j + shift]), v);
            return v;
          };
          auto gather_start = [&](casadi_int shift) {  // knots[start + shift]
            ccu_int v = cst(knots[smax + shift]);
            for (casadi_int j = smax; j-- > 0; ) v = sel(start_hit[k][j], cst(knots[j + shift]), v);
            return v;
          };
          // initial basis (nd_boor_eval): zeros; inside [knots[0], knots[end]]: x == knots[1] -> the first degree+1 ones,
          // else x == knots[end] -> boor[degree], else knots[L+degree] == x -> boor[degree-1], else boor[degree]
          Vals bo(2 * degree + 1, zero);
          {
            const ccu_int inside = op(OP_AND, op(OP_LE, cst(knots[0]), xk), op(OP_LE, xk, cst(knots[n_knots - 1])));
            const ccu_int c1 = op(OP_EQ, xk, cst(knots[1])), c2 = op(OP_EQ, xk, cst(knots[n_knots - 1]));
            const ccu_int c3 = op(OP_EQ, gather_L(degree), xk);
            const ccu_int n1 = op(OP_NOT, c1), n2 = op(OP_NOT, c2), n3 = op(OP_NOT, c3);
            const ccu_int only2 = op(OP_AND, n1, c2), only3 = op(OP_AND, op(OP_AND, n1, n2), c3), none = op(OP_AND, op(OP_AND, n1, n2), n3);
            for (casadi_int i = 0; i <= degree; ++i) {
              ccu_int v = c1;  // casadi_fill(boor, degree+1, 1.0)
              if (i == degree) v = op(OP_OR, v, op(OP_OR, only2, none));
              if (i == degree - 1) v = op(OP_OR, v, only3);
              bo[i] = op(OP_AND, inside, v);
            }
          }
          // casadi_de_boor(x, knots + start, 2*degree+2, degree, boor)
          const casadi_int nk = 2 * degree + 2;
          Vals kn(nk);
          for (casadi_int i = 0; i < nk; ++i) kn[i] = gather_start(i);
          for (casadi_int d = 1; d < degree + 1; ++d) {
            for (casadi_int i = 0; i < nk - d - 1; ++i) {
              ccu_int bv = zero;
              const ccu_int bottom = op(OP_SUB, kn[i + d], kn[i]);
              bv = sel(op(OP_NE, bottom, zero), op(OP_DIV, op(OP_MUL, op(OP_SUB, xk, kn[i]), bo[i]), bottom), bv);
              const ccu_int bottom2 = op(OP_SUB, kn[i + d + 1], kn[i + 1]);
              bv = sel(op(OP_NE, bottom2, zero),
                       op(OP_ADD, bv, op(OP_DIV, op(OP_MUL, op(OP_SUB, kn[i + d + 1], xk), bo[i + 1]), bottom2)), bv);
              bo[i] = bv;
            }
          }
          boor[k] = Vals(bo.begin(), bo.begin() + degree + 1);
          work *= static_cast<double>(degree + 1) * static_cast<double>(smax + 1);
        }
        casadi_assert(work <= 2e5, who + "the coefficient tensor is too large to be gathered by selects on the device ("
                      + str(work) + " selects per evaluation)");
        // casadi_tensor_ttv(ret, nd-1, nd, all_boor, boor_offset, starts, strides, c, m, 1.0, 0) on ret = 0
        Vals ret(m, zero);
        // coefficient c[off + j] with off = sum_k (start_k + i_k) * strides[k]: gathered over the one-hot starts, outer dimension first
        std::vector<casadi_int> idx(nd, 0);
        std::function<ccu_int(casadi_int, casadi_int, casadi_int)> gather = [&](casadi_int dim, casadi_int off, casadi_int j) -> ccu_int {
          if (dim < 0) return coeffs.at(off + j);
          const casadi_int smax = static_cast<casadi_int>(start_hit[dim].size()) - 1;
          ccu_int v = gather(dim - 1, off + (smax + idx[dim]) * bs->strides_[dim], j);
          for (casadi_int s2 = smax; s2-- > 0; ) v = sel(start_hit[dim][s2], gather(dim - 1, off + (s2 + idx[dim]) * bs->strides_[dim], j), v);
          return v;
        };
        std::function<void(casadi_int, ccu_int)> ttv = [&](casadi_int dim, ccu_int weight) {
          const casadi_int n_w = static_cast<casadi_int>(boor[dim].size());
          for (casadi_int i = 0; i < n_w; ++i) {
            idx[dim] = i;
            const ccu_int ww = op(OP_MUL, weight, boor[dim][i]);
            if (dim == 0) {
              for (casadi_int j = 0; j < m; ++j) ret[j] = op(OP_ADD, ret[j], op(OP_MUL, ww, gather(nd - 1, 0, j)));
            } else {
              ttv(dim - 1, ww);
            }
          }
        };
        ttv(nd - 1, one);
        return ret;
      }

      // A lookup table (interpolant(..., "linear", grid, values): casadi/solvers/linear_interpolant.cpp:85-96, 142-152) and its
      // Jacobian: casadi_interpn / casadi_interpn_grad (runtime/) replayed with their own operation order.  The left index
      // of every dimension (casadi_low) is data, and a tape has no indexed loads, so the grid points and the table entries a
      // lookup touches are gathered by bit-exact selects on one-hot flags "the index is j" -- comparisons of x with the
      // (constant) grid.  Cost: about a select per table entry and corner, so only tables of moderate size are accepted.
      void call_interpolant(const Function& f, const LinearInterpolant* I, bool grad, const std::vector<const Vals*>& arg,
                            std::vector<Vals*>& res) {
        const std::string who = "Map 'cuda': interpolant '" + f.name() + "': ";
        casadi_assert(I->batch_x_ == 1, who + "batch_x > 1 has no device lowering");
        const casadi_int ndim = I->ndim_, m = I->m_;
        const std::vector<casadi_int>& offset = I->offset_;
        // grid points and table entries: constants of the interpolant, or (the parametric variants, interpolant.hpp:108-146)
        // operands of the call -- one table per instance of the map
        const Vals* pgrid = I->has_parametric_grid() ? arg.at(I->arg_grid()) : nullptr;
        const Vals* pvals = I->has_parametric_values() ? arg.at(I->arg_values()) : nullptr;
        casadi_assert(!I->has_parametric_grid() || pgrid, who + "the grid operand is missing");
        casadi_assert(!I->has_parametric_values() || pvals, who + "the values operand is missing");
        auto gridv = [&](casadi_int e) { return pgrid ? pgrid->at(e) : cst(I->grid_.at(e)); };
        auto valv = [&](size_t e) { return pvals ? pvals->at(e) : cst(I->values_.at(e)); };
        double work = static_cast<double>(m) * (1 << ndim);
        for (casadi_int i = 0; i < ndim; ++i) work *= static_cast<double>(offset[i + 1] - offset[i]);
        casadi_assert(work <= 2e5, who + "the table is too large to be gathered by selects on the device (" + str(work) + " selects per lookup)");
        if (!res.at(0)) return;
        const ccu_int zero = cst(0.), one = cst(1.);
        auto sel = [&](ccu_int c, ccu_int a, ccu_int b2) {
          ccu_int h = lib.builder_select(b, c, a, b2);
          casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
          return h;
        };
        // casadi_interpn_weights: per dimension the one-hot flags of the left index j, alpha = (x - g[j]) / (g[j+1] - g[j])
        std::vector<std::vector<ccu_int>> hit(ndim);
        Vals alpha(ndim), delta(ndim);
        for (casadi_int i = 0; i < ndim; ++i) {
          const ccu_int xi = arg.at(0) ? arg[0]->at(i) : zero;
          const casadi_int ng = offset[i + 1] - offset[i];
          casadi_assert(ng >= 2, who + "a grid needs two points");
          Vals g(ng);
          for (casadi_int j = 0; j < ng; ++j) g[j] = gridv(offset[i] + j);
          hit[i].assign(ng - 1, zero);
          if (I->lookup_mode_.at(i) == 1) {
            // "exact": j = (casadi_int)((x - g0)*(ng-1)/dg) clamped to [0, ng-2]; trunc(t) >= k  <=>  t >= k for k >= 1
            const ccu_int t = op(OP_DIV, op(OP_MUL, op(OP_SUB, xi, g[0]), cst(static_cast<double>(ng - 1))), op(OP_SUB, g[ng - 1], g[0]));
            std::vector<ccu_int> ge(ng, zero);  // ge[k] = (t >= k), k = 1 .. ng-2
            for (casadi_int k = 1; k <= ng - 2; ++k) ge[k] = op(OP_LE, cst(static_cast<double>(k)), t);
            for (casadi_int j = 0; j <= ng - 2; ++j) {
              const ccu_int lo = j == 0 ? one : ge[j], hi = j == ng - 2 ? zero : ge[j + 1];
              hit[i][j] = op(OP_AND, lo, op(OP_NOT, hi));
            }
          } else {
            // linear / binary search over a strictly increasing grid: the first j in [0, ng-2) with x < g[j+1], else ng-2
            std::vector<ccu_int> lt(ng, zero);  // lt[k] = (x < g[k]), k = 1 .. ng-2
            for (casadi_int k = 1; k <= ng - 2; ++k) lt[k] = op(OP_LT, xi, g[k]);
            for (casadi_int j = 0; j <= ng - 2; ++j) {
              const ccu_int below = j == ng - 2 ? one : lt[j + 1], not_earlier = j == 0 ? one : op(OP_NOT, lt[j]);
              hit[i][j] = op(OP_AND, below, not_earlier);
            }
          }
          ccu_int gj = g[ng - 2], gj1 = g[ng - 1];
          for (casadi_int j = ng - 2; j-- > 0; ) { gj = sel(hit[i][j], g[j], gj); gj1 = sel(hit[i][j], g[j + 1], gj1); }
          delta[i] = op(OP_SUB, gj1, gj);
          alpha[i] = op(OP_DIV, op(OP_SUB, xi, gj), delta[i]);
        }
        // the table entry values[(index + corner) . strides * m + k], gathered dimension by dimension
        std::vector<casadi_int> ngs(ndim), stride(ndim);
        for (casadi_int i = 0, ld = 1; i < ndim; ++i) { ngs[i] = offset[i + 1] - offset[i]; stride[i] = ld; ld *= ngs[i]; }
        auto gather = [&](const std::vector<casadi_int>& corner, casadi_int k) {
          // level d holds, for every multi-index of the dimensions >= d, the entry selected in the dimensions < d
          std::vector<ccu_int> cur(static_cast<size_t>(stride[ndim - 1] * ngs[ndim - 1]));
          for (size_t e = 0; e < cur.size(); ++e) cur[e] = valv(e * m + k);
          casadi_int inner = 1;  // entries per block of the dimensions already resolved (always 1 after resolution)
          casadi_int count = static_cast<casadi_int>(cur.size());
          for (casadi_int d = 0; d < ndim; ++d) {
            const casadi_int ng = ngs[d], blocks = count / ng;
            std::vector<ccu_int> next(static_cast<size_t>(blocks));
            for (casadi_int bl = 0; bl < blocks; ++bl) {
              // entries bl*ng + (j + corner[d]), j = 0 .. ng-2
              ccu_int v = cur[static_cast<size_t>(bl * ng + (ng - 2) + corner[d])];
              for (casadi_int j = ng - 2; j-- > 0; ) v = sel(hit[d][j], cur[static_cast<size_t>(bl * ng + j + corner[d])], v);
              next[static_cast<size_t>(bl)] = v;
            }
            cur.swap(next);
            count = blocks;
          }
          (void)inner;
          return cur.at(0);
        };
        std::vector<casadi_int> corner(ndim, 0);
        auto flip = [&]() {  // casadi_flip
          for (casadi_int i = 0; i < ndim; ++i) { if (corner[i]) corner[i] = 0; else { corner[i] = 1; return true; } }
          return false;
        };
        if (!grad) {
          // casadi_interpn: res = 0; per corner res[k] += c * value, c = prod_i (corner_i ? alpha_i : 1 - alpha_i) from c = 1
          Vals r(m, zero);
          do {
            ccu_int c = one;
            for (casadi_int i = 0; i < ndim; ++i) c = op(OP_MUL, c, corner[i] ? alpha[i] : op(OP_SUB, one, alpha[i]));
            for (casadi_int k = 0; k < m; ++k) r[k] = op(OP_ADD, r[k], op(OP_MUL, c, gather(corner, k)));
          } while (flip());
          *res[0] = r;
        } else {
          // casadi_interpn_grad: per corner v = value (coeff[i] = the partial product before dimension i), then from the last
          // dimension down grad[i] +/-= v * coeff[i], v *= (alpha_i | 1 - alpha_i); finally grad[i] /= g[j+1] - g[j]
          Vals gr(ndim * m, zero);
          do {
            Vals coeff(ndim);
            ccu_int c = one;
            for (casadi_int i = 0; i < ndim; ++i) {
              coeff[i] = c;
              c = op(OP_MUL, c, corner[i] ? alpha[i] : op(OP_SUB, one, alpha[i]));
            }
            Vals v(m);
            for (casadi_int k = 0; k < m; ++k) v[k] = op(OP_ADD, zero, gather(corner, k));  // (casadi_clear(v); v[k] += values[k])
            for (casadi_int i = ndim; i-- > 0; ) {
              for (casadi_int k = 0; k < m; ++k) {
                const ccu_int term = op(OP_MUL, v[k], coeff[i]);
                gr[i * m + k] = corner[i] ? op(OP_ADD, gr[i * m + k], term) : op(OP_SUB, gr[i * m + k], term);
                v[k] = op(OP_MUL, v[k], corner[i] ? alpha[i] : op(OP_SUB, one, alpha[i]));
              }
            }
          } while (flip());
          for (casadi_int i = 0; i < ndim; ++i)
            for (casadi_int k = 0; k < m; ++k) gr[i * m + k] = op(OP_DIV, gr[i * m + k], delta[i]);
          casadi_assert(static_cast<casadi_int>(res[0]->size()) == ndim * m, who + "unexpected Jacobian size");
          *res[0] = gr;
        }
      }

      // casadi_project on handles: the nonzeros of pattern `to` taken from a matrix with pattern `from`, zero where absent
      Vals project_vals(const Vals& v, const Sparsity& from, const Sparsity& to) {
        if (from == to) return v;
        casadi_assert(from.size() == to.size(), "Map 'cuda': projection between patterns of different dimension");
        Vals r(to.nnz(), cst(0.));
        const casadi_int *fc = from.colind(), *fr = from.row(), *tc = to.colind(), *tr = to.row();
        for (casadi_int c = 0; c < to.size2(); ++c) {
          casadi_int kf = fc[c];
          for (casadi_int kt = tc[c]; kt < tc[c + 1]; ++kt) {
            while (kf < fc[c + 1] && fr[kf] < tr[kt]) ++kf;
            if (kf < fc[c + 1] && fr[kf] == tr[kt]) r[kt] = v.at(kf);
          }
        }
        return r;
      }

      // Switch::eval (switch.cpp:153-211; Function::conditional / if_else): the reference evaluates the one case the index
      // selects -- k = static_cast<casadi_int>(index), the default for k outside [0, n) -- with operands and results
      // projected between the sparsity of the switch and of the case.  A thread of a map cannot branch per instance, so
      // every case is evaluated and the results are merged with bit-exact selects on (trunc(index) == k).
      void call_switch(const Function& f, const Switch* sw, const std::vector<const Vals*>& arg, std::vector<Vals*>& res) {
        const casadi_int nf = static_cast<casadi_int>(sw->f_.size()), n_in = f.n_in(), n_out = f.n_out();
        casadi_assert(!sw->f_def_.is_null(), "Map 'cuda': switch '" + f.name() + "' has no default case");
        const ccu_int zero = cst(0.);
        const ccu_int index = arg.at(0) && !arg[0]->empty() ? arg[0]->at(0) : zero;
        auto sel = [&](ccu_int c, ccu_int a, ccu_int b2) {
          ccu_int h = lib.builder_select(b, c, a, b2);
          casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
          return h;
        };
        // truncation toward zero of the conversion to an integer
        const ccu_int trunc = sel(op(OP_LE, zero, index), op(OP_FLOOR, index), op(OP_CEIL, index));
        auto eval_case = [&](const Function& fk, std::vector<Vals>& out) {
          casadi_assert(!fk.is_null(), "Map 'cuda': switch '" + f.name() + "' has an empty case");
          std::vector<Vals> a(n_in - 1), r(n_out);
          std::vector<const Vals*> ap(n_in - 1, nullptr);
          std::vector<Vals*> rp(n_out, nullptr);
          for (casadi_int i = 0; i + 1 < n_in; ++i) {
            if (!arg.at(i + 1)) continue;
            a[i] = project_vals(*arg[i + 1], f.sparsity_in(i + 1), fk.sparsity_in(i));
            ap[i] = &a[i];
          }
          for (casadi_int i = 0; i < n_out; ++i) {
            if (!res.at(i)) continue;
            r[i].assign(fk.nnz_out(i), zero);
            rp[i] = &r[i];
          }
          call(fk, ap, rp);
          out.resize(n_out);
          for (casadi_int i = 0; i < n_out; ++i)
            if (res[i]) out[i] = project_vals(r[i], fk.sparsity_out(i), f.sparsity_out(i));
        };
        std::vector<Vals> acc;
        eval_case(sw->f_def_, acc);
        for (casadi_int k = 0; k < nf; ++k) {
          std::vector<Vals> rk;
          eval_case(sw->f_[k], rk);
          const ccu_int hit = op(OP_EQ, trunc, cst(static_cast<double>(k)));
          for (casadi_int i = 0; i < n_out; ++i)
            if (res[i]) for (size_t e = 0; e < acc[i].size(); ++e) acc[i][e] = sel(hit, rk[i][e], acc[i][e]);
        }
        for (casadi_int i = 0; i < n_out; ++i) if (res[i]) *res[i] = acc[i];
      }

      void call(const Function& f, const std::vector<const Vals*>& arg, std::vector<Vals*>& res) {
        if (f.is_a("SXFunction")) {
          call_sx(f, arg, res);
        } else if (f.is_a("MXFunction")) {
          call_mx(f, arg, res);
        } else if (auto* I = dynamic_cast<const FixedStepIntegrator*>(f.get())) {
          call_fixed_step(f, I, arg, res);
        } else if (f.is_a("Map", true)) {
          // a map embedded in the function (g.map(k) called from MX; any parallelization evaluates like "serial"):
          // Map::eval_gen (map.cpp:141-157), instance i reads arg[j] + i*nnz_in(j) and writes res[j] + i*nnz_out(j)
          Dict inf = f.info();
          call_repeated(inf.at("f").to_function(), inf.at("n").to_int(), std::vector<bool>(f.n_in(), false),
                        std::vector<bool>(f.n_out(), false), arg, res);
        } else if (auto* sw = dynamic_cast<const Switch*>(f.get())) {
          call_switch(f, sw, arg, res);
        } else if (f.class_name() == "BSplineInterpolant") {
          // BSplineInterpolant::eval (casadi/solvers/bspline_interpolant.cpp:185-190) evaluates its MX function S_, one B-spline node
          call(static_cast<const BSplineInterpolant*>(f.get())->S_, arg, res);
        } else if (f.class_name() == "LinearInterpolant") {
          call_interpolant(f, static_cast<const LinearInterpolant*>(f.get()), false, arg, res);
        } else if (f.class_name() == "LinearInterpolantJac") {
          const Function& of = f->derivative_of_;
          casadi_assert(!of.is_null() && of.class_name() == "LinearInterpolant", "Map 'cuda': '" + f.name() + "' has no interpolant");
          call_interpolant(f, static_cast<const LinearInterpolant*>(of.get()), true, arg, res);
        } else if (auto* ms = dynamic_cast<const MapSum*>(f.get())) {
          // MapSum::eval_gen (mapsum.cpp:154-186): reduced inputs are shared, reduced outputs are cleared and then
          // accumulated instance by instance, in index order (casadi_add: y += x)
          struct Peek : public MapSum { using MapSum::f_; using MapSum::n_; using MapSum::reduce_in_; using MapSum::reduce_out_; };
          Function MapSum::* pf = &Peek::f_;
          casadi_int MapSum::* pn = &Peek::n_;
          std::vector<bool> MapSum::* pri = &Peek::reduce_in_;
          std::vector<bool> MapSum::* pro = &Peek::reduce_out_;
          call_repeated(ms->*pf, ms->*pn, ms->*pri, ms->*pro, arg, res);
        } else {
          casadi_error("Map 'cuda': embedded function '" + f.name() + "' of class " + f.class_name()
                       + " has no device lowering");
        }
      }
    };
  } // namespace

  CudaMap::CudaMap(const std::string& name, const Function& f, casadi_int n)
    : Map(name, f, n), rep_(1), flatten_(true), device_(0), builder_(nullptr), has_flag_(false), newton_(false) {
  }

  CudaMap::CudaMap(DeserializingStream& s) : Map(s), rep_(1), flatten_(true), device_(0), builder_(nullptr), has_flag_(false), newton_(false) {
    s.unpack("CudaMap::in_groups", in_groups_);
    s.unpack("CudaMap::out_groups", out_groups_);
    s.unpack("CudaMap::flatten", flatten_);
    // The device program is not serialized (Map::serialize_body packs f_ and n_ only, map.cpp:94-98):
    // it is re-exported from f_, exactly like a freshly created map
    export_function();
  }

  void CudaMap::serialize_body(SerializingStream &s) const {
    Map::serialize_body(s);
    s.pack("CudaMap::in_groups", in_groups_);
    s.pack("CudaMap::out_groups", out_groups_);
    s.pack("CudaMap::flatten", flatten_);
  }

  // df.map(n, "cuda") with piece-major derivative blocks; `first` = index of the first seed input of df
  static Function grouped_derivative_map(const Function& df, casadi_int n, casadi_int ndir, casadi_int first_seed,
                                         casadi_int n_seed, casadi_int n_sens, const std::string& name,
                                         const std::vector<std::string>& inames, const std::vector<std::string>& onames,
                                         const Dict& opts) {
    CudaMap* cm = new CudaMap("cudamap" + str(n) + "_" + df.name(), df, n);
    std::vector<casadi_int> gi(df.n_in(), 1), go(df.n_out(), 1);
    for (casadi_int i = 0; i < n_seed; ++i) gi.at(first_seed + i) = ndir;
    for (casadi_int i = 0; i < n_sens; ++i) go.at(i) = ndir;
    cm->set_groups(gi, go);
    Function dm = Function::create(cm, Dict());
    std::vector<MX> arg = dm.mx_in();
    std::vector<MX> res = dm(arg);
    Dict options = opts;
    options["allow_duplicate_io_names"] = true;
    return Function(name, arg, res, inames, onames, options);
  }

  Function CudaMap::get_forward(casadi_int nfwd, const std::string& name, const std::vector<std::string>& inames,
                                const std::vector<std::string>& onames, const Dict& opts) const {
    // one direction needs no permutation; nested (flattened) and already grouped maps take the reference's route
    if (nfwd <= 1 || rep_ != 1 || !in_groups_.empty() || !out_groups_.empty())
      return Map::get_forward(nfwd, name, inames, onames, opts);
    // df(arg..., res..., fseed...) -> fsens...   (function_internal.cpp: forward(nfwd) signature)
    return grouped_derivative_map(f_.forward(nfwd), n_, nfwd, n_in_ + n_out_, n_in_, n_out_, name, inames, onames, opts);
  }

  Function CudaMap::get_reverse(casadi_int nadj, const std::string& name, const std::vector<std::string>& inames,
                                const std::vector<std::string>& onames, const Dict& opts) const {
    if (nadj <= 1 || rep_ != 1 || !in_groups_.empty() || !out_groups_.empty())
      return Map::get_reverse(nadj, name, inames, onames, opts);
    // df(arg..., res..., aseed...) -> asens...
    return grouped_derivative_map(f_.reverse(nadj), n_, nadj, n_in_ + n_out_, n_out_, n_in_, name, inames, onames, opts);
  }

  CudaMap::~CudaMap() {
    clear_mem();
    if (builder_) cuda_lib().builder_destroy(builder_);
  }

  bool CudaMap::is_a(const std::string& type, bool recursive) const {
    return type=="CudaMap"
      || (recursive && Map::is_a(type, recursive));
  }

  CudaMap::Tape CudaMap::export_tape(const Function& f) {
    casadi_assert(f.is_a("SXFunction"), "Tape export needs an SXFunction, got " + f.class_name());
    // enum Operation values with non-slot operands (calculus.hpp:60-218)
    const int op_const = OP_CONST, op_input = OP_INPUT, op_output = OP_OUTPUT;
    Tape t;
    casadi_int n = f.n_instructions();
    t.op.resize(n); t.i0.assign(n, 0); t.i1.assign(n, 0); t.i2.assign(n, 0); t.d.assign(n, 0.);
    for (casadi_int k=0; k<n; ++k) {
      int op = static_cast<int>(f.instruction_id(k));
      t.op[k] = op;
      std::vector<casadi_int> in = f.instruction_input(k), out = f.instruction_output(k);
      if (op==op_const) {
        t.i0[k] = static_cast<int>(out.at(0));
        t.d[k] = f.instruction_constant(k);
      } else if (op==op_input) {
        t.i0[k] = static_cast<int>(out.at(0));
        t.i1[k] = static_cast<int>(in.at(0));
        t.i2[k] = static_cast<int>(in.at(1));
      } else if (op==op_output) {
        t.i0[k] = static_cast<int>(out.at(0));
        t.i2[k] = static_cast<int>(out.at(1));
        t.i1[k] = static_cast<int>(in.at(0));
      } else if (op==OP_CALL) {
        casadi_error("Map 'cuda': function '" + f.name() + "' embeds a function call (OP_CALL); "
                     "inline it (Function::expand or the 'never_inline'/'always_inline' options) first");
      } else {
        t.i0[k] = static_cast<int>(out.at(0));
        t.i1[k] = static_cast<int>(in.at(0));
        t.i2[k] = static_cast<int>(in.size()>1 ? in.at(1) : in.at(0));
      }
    }
    t.sz_w = f.sz_w();
    for (casadi_int j=0; j<f.n_in(); ++j) t.nnz_in.push_back(f.nnz_in(j));
    for (casadi_int j=0; j<f.n_out(); ++j) t.nnz_out.push_back(f.nnz_out(j));
    return t;
  }

  static bool lowerable_class(const Function& f);

  void CudaMap::export_function() {
    builder_ = nullptr;
    has_flag_ = false;
    // Nested maps are flattened: Function::map(n, par, max_num_threads) (function.cpp:829-858) builds
    // f.map(d, "serial").map(T, par), whose memory layout -- T blocks of d consecutive instances -- is exactly that
    // of f.map(d*T).  Mapping the leaf over d*T device threads keeps the tape short and the parallelism full
    // (expanding the inner map instead would give T threads a d times longer tape).
    leaf_ = f_;
    rep_ = 1;
    while (flatten_ && leaf_.is_a("Map", true)) {
      Dict inf = leaf_.info();
      rep_ *= inf.at("n").to_int();
      leaf_ = inf.at("f").to_function();
    }
    newton_ = is_newton(leaf_);
    if (newton_) {
      // data-dependent iteration: two tapes driven from the host over device-resident state (NewtonPlan)
      newton_plan_ = newton_plan(leaf_);
      tape_ = Tape();
      for (casadi_int j=0; j<leaf_.n_in(); ++j) tape_.nnz_in.push_back(leaf_.nnz_in(j));
      for (casadi_int j=0; j<leaf_.n_out(); ++j) tape_.nnz_out.push_back(leaf_.nnz_out(j));
      return;
    }
    if (leaf_.is_a("SXFunction")) {
      sx_ = leaf_;
    } else {
      // An MX function whose nodes all have an SX evaluation (mapaccum/fold towers, wrapped maps)
      // collapses to one SX tape ...
      // (a wrapped non-MX function "expands" to an SX function that still calls it: OP_CALL is not an expansion)
      bool expanded = false;
      try {
        sx_ = leaf_.expand();
        expanded = true;
        for (casadi_int k = 0; k < sx_.n_instructions() && expanded; ++k) expanded = sx_.instruction_id(k) != OP_CALL;
      } catch (std::exception& e) {
        expanded = false;
      }
      if (!expanded) {
        // ... anything else (a Linsol call, solve_impl.hpp:57-73: "eval_sx not defined"; a fixed-step integrator) is
        // lowered node by node through the tape builder; unsupported nodes raise from there
        casadi_assert(lowerable_class(leaf_), "Map 'cuda': function '" + leaf_.name() + "' (" + leaf_.class_name()
                      + ") is neither an SX function, an MX function nor a fixed-step integrator");
        lower_mx();
        return;
      }
    }
    casadi_assert(!sx_.has_free(), "Map 'cuda': function '" + leaf_.name() + "' has free variables "
                  + str(sx_.get_free()) + " and cannot be evaluated");
    tape_ = export_tape(sx_);
  }

  // Lower `leaf` node by node through the tape builder; returns the builder (owned by the caller)
  static void* lower_function(const Function& leaf, bool* has_flag) {
    CudaLib& lib = cuda_lib();
    casadi_assert(lib.handle!=nullptr, "Map 'cuda': " + lib.error);
    casadi_assert(!leaf.has_free(), "Map 'cuda': function '" + leaf.name() + "' has free variables "
                  + str(leaf.get_free()) + " and cannot be evaluated");
    Lowering L(lib);
    *has_flag = false;
    try {
      std::vector<Vals> in(leaf.n_in()), out(leaf.n_out());
      std::vector<const Vals*> a(leaf.n_in());
      std::vector<Vals*> r(leaf.n_out());
      for (casadi_int j=0; j<leaf.n_in(); ++j) {
        in[j].resize(leaf.nnz_in(j));
        for (casadi_int e=0; e<leaf.nnz_in(j); ++e) in[j][e] = lib.builder_input(L.b, j, e);
        a[j] = &in[j];
      }
      for (casadi_int j=0; j<leaf.n_out(); ++j) {
        out[j].assign(leaf.nnz_out(j), L.cst(0.));
        r[j] = &out[j];
      }
      L.call(leaf, a, r);
      for (casadi_int j=0; j<leaf.n_out(); ++j)
        for (casadi_int e=0; e<leaf.nnz_out(j); ++e) lib.builder_output(L.b, j, e, out[j][e]);
      // instances whose QR factorisation is numerically singular make the reference's map fail
      // (LinsolQr::nfact returns 1, linsol_qr.cpp:146-163): counted in one extra, summed output
      if (L.fail_count >= 0) {
        lib.builder_output(L.b, leaf.n_out(), 0, L.fail_count);
        *has_flag = true;
      }
    } catch (...) {
      lib.builder_destroy(L.b);
      throw;
    }
    return L.b;
  }

  static bool lowerable_class(const Function& f) {
    return f.is_a("MXFunction") || dynamic_cast<const FixedStepIntegrator*>(f.get()) != nullptr;
  }

  void CudaMap::lower_mx() {
    builder_ = lower_function(leaf_, &has_flag_);
    tape_ = Tape();
    for (casadi_int j=0; j<leaf_.n_in(); ++j) tape_.nnz_in.push_back(leaf_.nnz_in(j));
    for (casadi_int j=0; j<leaf_.n_out(); ++j) tape_.nnz_out.push_back(leaf_.nnz_out(j));
    if (has_flag_) tape_.nnz_out.push_back(1);
  }

  CudaMap::Tape CudaMap::lowered_tape(const Function& f) {
    if (f.is_a("SXFunction")) return export_tape(f);
    try {
      return export_tape(f.expand());
    } catch (std::exception& e) {  // not expandable, or the expansion still holds a call (export_tape refuses OP_CALL)
      casadi_assert(lowerable_class(f), "Map 'cuda': function '" + f.name() + "' (" + f.class_name()
                    + ") is neither an SX function, an MX function nor a fixed-step integrator");
    }
    CudaLib& lib = cuda_lib();
    bool flag = false;
    void* b = lower_function(f, &flag);
    Tape t;
    ccu_int sz_w = 0;
    ccu_int n = lib.builder_export(b, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &sz_w);
    t.op.resize(n); t.i0.resize(n); t.i1.resize(n); t.i2.resize(n); t.d.resize(n);
    lib.builder_export(b, get_ptr(t.op), get_ptr(t.i0), get_ptr(t.i1), get_ptr(t.i2), get_ptr(t.d), n, &sz_w);
    lib.builder_destroy(b);
    t.sz_w = sz_w;
    for (casadi_int j=0; j<f.n_in(); ++j) t.nnz_in.push_back(f.nnz_in(j));
    for (casadi_int j=0; j<f.n_out(); ++j) t.nnz_out.push_back(f.nnz_out(j));
    if (flag) t.nnz_out.push_back(1);
    return t;
  }

  // ------------------------------------------------------------------------------------------------
  // Newton rootfinder under the map (see NewtonPlan in cuda_map.hpp)
  // ------------------------------------------------------------------------------------------------
  namespace {
    // The option members of the plugin class are protected; a pointer to member formed in a derived scope has the type
    // `T Newton::*` and reads them from any Newton (no object of this type ever exists, nothing of the plugin is linked)
    struct NewtonOptions : public Newton {
      static casadi_int max_iter(const Newton* p) { return p->*(&NewtonOptions::max_iter_); }
      static double abstol(const Newton* p) { return p->*(&NewtonOptions::abstol_); }
      static double abstol_step(const Newton* p) { return p->*(&NewtonOptions::abstolStep_); }
      static bool line_search(const Newton* p) { return p->*(&NewtonOptions::line_search_); }
    };

    struct FastNewtonOptions : public FastNewton {
      static casadi_int max_iter(const FastNewton* p) { return p->*(&FastNewtonOptions::max_iter_); }
      static double abstol(const FastNewton* p) { return p->*(&FastNewtonOptions::abstol_); }
      static double abstol_step(const FastNewton* p) { return p->*(&FastNewtonOptions::abstolStep_); }
    };

    CudaMap::Tape export_builder(CudaLib& lib, void* b, const std::vector<casadi_int>& nnz_in,
                                 const std::vector<casadi_int>& nnz_out) {
      CudaMap::Tape t;
      ccu_int sz_w = 0;
      ccu_int n = lib.builder_export(b, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &sz_w);
      t.op.resize(n); t.i0.resize(n); t.i1.resize(n); t.i2.resize(n); t.d.resize(n);
      lib.builder_export(b, get_ptr(t.op), get_ptr(t.i0), get_ptr(t.i1), get_ptr(t.i2), get_ptr(t.d), n, &sz_w);
      t.sz_w = sz_w;
      t.nnz_in = nnz_in;
      t.nnz_out = nnz_out;
      return t;
    }
  } // namespace

  bool CudaMap::is_newton(const Function& f) {
    return (f.class_name() == "Newton" || f.class_name() == "FastNewton") && dynamic_cast<const Rootfinder*>(f.get()) != nullptr;
  }

  CudaMap::NewtonPlan CudaMap::newton_plan(const Function& rf) {
    casadi_assert(is_newton(rf), "Map 'cuda': '" + rf.name() + "' is not a Newton rootfinder");
    CudaLib& lib = cuda_lib();
    casadi_assert(lib.handle!=nullptr, "Map 'cuda': " + lib.error);
    const Rootfinder* R = dynamic_cast<const Rootfinder*>(rf.get());
    // "fast_newton" (casadi/solvers/fast_newton.cpp:140-166 + runtime/casadi_newton.hpp): the same iteration without a line
    // search, with casadi_qr as its built-in linear solver, norms by casadi_norm_inf (fmax), tolerances that only count when
    // positive, and the step applied BEFORE it is tested
    const bool fast = rf.class_name() == "FastNewton";
    NewtonPlan P;
    P.name = rf.name();
    P.n = R->n_; P.iin = R->iin_; P.iout = R->iout_;
    P.max_iter = fast ? FastNewtonOptions::max_iter(static_cast<const FastNewton*>(R)) : NewtonOptions::max_iter(static_cast<const Newton*>(R));
    P.line_search = fast ? false : NewtonOptions::line_search(static_cast<const Newton*>(R));
    P.error_on_fail = R->error_on_fail_;
    const double abstol = fast ? FastNewtonOptions::abstol(static_cast<const FastNewton*>(R)) : NewtonOptions::abstol(static_cast<const Newton*>(R));
    const double abstol_step = fast ? FastNewtonOptions::abstol_step(static_cast<const FastNewton*>(R))
                                    : NewtonOptions::abstol_step(static_cast<const Newton*>(R));
    const double inf = std::numeric_limits<double>::infinity();
    casadi_assert(P.max_iter >= 1, "Map 'cuda': Newton rootfinder '" + rf.name() + "' with max_iter < 1");
    const casadi_int n = P.n, n_in = rf.n_in(), n_out = rf.n_out();
    for (casadi_int j = 0; j < n_in; ++j) P.nnz_in.push_back(rf.nnz_in(j));
    for (casadi_int j = 0; j < n_out; ++j) { P.nnz_out.push_back(rf.nnz_out(j)); if (j != P.iout) P.aux.push_back(j); }
    casadi_assert(rf.nnz_in(P.iin) == n && rf.nnz_out(P.iout) == n, "Map 'cuda': Newton rootfinder with a sparse unknown");
    const Function& jac = R->get_function("jac_g_x");
    // (fast_newton has no separate residual function; the line-search tape is never launched for it, the oracle fills its place)
    const Function g = fast ? R->oracle() : R->get_function("g");
    // tape signature (both tapes)
    std::vector<casadi_int> t_in = P.nnz_in, t_out;
    t_in.push_back(n); t_in.push_back(NEWTON_SC);
    for (casadi_int j : P.aux) t_in.push_back(P.nnz_out[j]);
    t_out.push_back(n); t_out.push_back(n); t_out.push_back(NEWTON_SC);
    for (casadi_int j : P.aux) t_out.push_back(P.nnz_out[j]);
    t_out.push_back(NEWTON_COUNTS);
    for (int which = 0; which < 2; ++which) {
      Lowering L(lib);
      try {
        auto sel = [&](ccu_int c, ccu_int a, ccu_int b2) {
          ccu_int h = lib.builder_select(L.b, c, a, b2);
          casadi_assert(h >= 0, "Map 'cuda': " + std::string(lib.last_error()));
          return h;
        };
        std::vector<Vals> in(t_in.size());
        for (size_t j = 0; j < t_in.size(); ++j) {
          in[j].resize(t_in[j]);
          for (casadi_int e = 0; e < t_in[j]; ++e) in[j][e] = lib.builder_input(L.b, static_cast<ccu_int>(j), e);
        }
        const Vals& X = in[P.iin];
        const Vals& DX = in[n_in];
        const Vals& SC = in[n_in + 1];
        const ccu_int zero = L.cst(0.), one = L.cst(1.);
        const ccu_int sc_abstol = SC[0], sc_step = SC[1], sc_alpha = SC[2], sc_active = SC[3], sc_ls = SC[4], sc_failed = SC[5];
        Vals Xn(n), DXn(n), SCn(NEWTON_SC), CNT(NEWTON_COUNTS);
        std::vector<Vals> aux_new(P.aux.size());
        if (which == 0) {
          // ---- Newton direction (newton.cpp:152-196): jac_g_x at X, max|F| test, factorise + solve, step test
          std::vector<const Vals*> a(n_in, nullptr);
          for (casadi_int j = 0; j < n_in; ++j) a[j] = &in[j];
          std::vector<Vals> r(jac.n_out());
          std::vector<Vals*> rp(jac.n_out(), nullptr);
          for (casadi_int j = 0; j < jac.n_out(); ++j) { r[j].assign(jac.nnz_out(j), zero); rp[j] = &r[j]; }
          L.call(jac, a, rp);
          const Vals& J = r[0];
          const Vals& F = r[1 + P.iout];
          ccu_int abst = zero, conv1 = zero;
          if (fast) {
            if (abstol > 0) {  // casadi_norm_inf(n, g) <= abstol
              for (casadi_int i = 0; i < n; ++i) abst = L.op(OP_FMAX, abst, L.op(OP_FABS, F[i]));
              conv1 = L.op(OP_LE, abst, L.cst(abstol));
            }
          } else if (abstol != inf) {
            for (casadi_int i = 0; i < n; ++i) {  // abstol = std::max(abstol, fabs(f[i])): (a < b) ? b : a
              ccu_int fa = L.op(OP_FABS, F[i]);
              abst = sel(L.op(OP_LT, abst, fa), fa, abst);
            }
            conv1 = L.op(OP_LE, abst, L.cst(abstol));
          }
          Vals dx = F;
          if (fast) {
            // casadi_qr + casadi_qr_solve on the Jacobian's pattern (FastNewton::init: sp_jac_.qr_sparse, fast_newton.cpp:100)
            const Sparsity& spj = jac.sparsity_out(0);
            Sparsity spv, spr;
            std::vector<casadi_int> prinv, pc;
            spj.qr_sparse(spv, spr, prinv, pc);
            std::vector<ccu_int> a1 = Lowering::pattern(spj), v1 = Lowering::pattern(spv), r1 = Lowering::pattern(spr),
                                 pi(prinv.begin(), prinv.end()), pcc(pc.begin(), pc.end());
            ccu_int nullity = -1;
            casadi_assert(lib.builder_qr(L.b, a1.data(), v1.data(), r1.data(), pi.data(), pcc.data(), J.data(), dx.data(), 1, 0, 1e-12,
                                         &nullity) == 0, "Map 'cuda': " + std::string(lib.last_error()));
          } else {
            L.linsol_solve(R->linsol_, J, dx, 1, false);
          }
          const ccu_int singular = L.fail_count >= 0 ? L.fail_count : zero;
          L.fail_count = -1;
          ccu_int st = zero, conv2 = zero;
          if (fast) {
            if (abstol_step > 0) {
              for (casadi_int i = 0; i < n; ++i) st = L.op(OP_FMAX, st, L.op(OP_FABS, dx[i]));
              conv2 = L.op(OP_LE, st, L.cst(abstol_step));
            }
          } else if (abstol_step != inf) {
            for (casadi_int i = 0; i < n; ++i) {
              ccu_int fa = L.op(OP_FABS, dx[i]);
              st = sel(L.op(OP_LT, st, fa), fa, st);
            }
            conv2 = L.op(OP_LE, st, L.cst(abstol_step));
          }
          const ccu_int act = sc_active;
          const ccu_int cont = L.op(OP_AND, act, L.op(OP_NOT, L.op(OP_OR, conv1, conv2)));  // iterates on after this phase
          const ccu_int upd = L.op(OP_AND, act, L.op(OP_NOT, conv1));                        // reached the linear solve
          const ccu_int minus_one = L.cst(-1.);
          for (casadi_int i = 0; i < n; ++i) {
            // without line search: casadi_axpy(n, -alpha, f, x) with alpha = 1
            // (fast_newton applies the step whenever the residual test did not stop it, also a step that then ends the iteration)
            Xn[i] = P.line_search ? X[i] : sel(fast ? upd : cont, L.op(OP_ADD, X[i], L.op(OP_MUL, minus_one, dx[i])), X[i]);
            DXn[i] = sel(upd, dx[i], DX[i]);
          }
          SCn[0] = sel(act, abst, sc_abstol);
          SCn[1] = sel(upd, st, sc_step);
          SCn[2] = sel(cont, one, sc_alpha);
          SCn[3] = cont;
          SCn[4] = P.line_search ? cont : zero;
          SCn[5] = sc_failed;
          for (size_t k = 0; k < P.aux.size(); ++k) {
            const Vals& old = in[n_in + 2 + k];
            const Vals& nw = r[1 + P.aux[k]];
            aux_new[k].resize(old.size());
            for (size_t e = 0; e < old.size(); ++e) aux_new[k][e] = sel(act, nw[e], old[e]);
          }
          CNT[0] = SCn[3]; CNT[1] = SCn[4]; CNT[2] = L.op(OP_AND, upd, singular); CNT[3] = SCn[5];
        } else {
          // ---- one line-search trial (newton.cpp:199-221): x_trial = x - alpha*dx, g, acceptance, give up or halve alpha
          Vals xt(n);
          const ccu_int na = L.op(OP_NEG, sc_alpha);
          for (casadi_int i = 0; i < n; ++i) xt[i] = L.op(OP_ADD, X[i], L.op(OP_MUL, na, DX[i]));
          std::vector<const Vals*> a(n_in, nullptr);
          for (casadi_int j = 0; j < n_in; ++j) a[j] = j == P.iin ? &xt : &in[j];
          std::vector<Vals> r(g.n_out());
          std::vector<Vals*> rp(g.n_out(), nullptr);
          for (casadi_int j = 0; j < g.n_out(); ++j) { r[j].assign(g.nnz_out(j), zero); rp[j] = &r[j]; }
          L.call(g, a, rp);
          casadi_assert(L.fail_count < 0, "Map 'cuda': linear solves inside the residual function of a rootfinder");
          const Vals& Ft = r[P.iout];
          ccu_int nt = zero;  // casadi_norm_inf: fmax(ret, fabs(x))
          for (casadi_int i = 0; i < n; ++i) nt = L.op(OP_FMAX, nt, L.op(OP_FABS, Ft[i]));
          const ccu_int thr = L.op(OP_MUL, L.op(OP_SUB, one, L.op(OP_DIV, sc_alpha, L.cst(2.))), sc_abstol);
          const ccu_int accept = L.op(OP_LE, nt, thr);
          const ccu_int limit = L.op(OP_LE, L.op(OP_MUL, sc_alpha, sc_step), L.cst(abstol_step));
          const ccu_int giveup = L.op(OP_AND, L.op(OP_NOT, accept), limit);
          const ccu_int acc = L.op(OP_AND, sc_ls, accept), gv = L.op(OP_AND, sc_ls, giveup);
          const ccu_int ls_next = L.op(OP_AND, sc_ls, L.op(OP_AND, L.op(OP_NOT, accept), L.op(OP_NOT, giveup)));
          for (casadi_int i = 0; i < n; ++i) { Xn[i] = sel(acc, xt[i], X[i]); DXn[i] = DX[i]; }
          SCn[0] = sc_abstol;
          SCn[1] = sc_step;
          SCn[2] = sel(ls_next, L.op(OP_MUL, sc_alpha, L.cst(0.5)), sc_alpha);
          SCn[3] = L.op(OP_AND, sc_active, L.op(OP_NOT, gv));
          SCn[4] = ls_next;
          SCn[5] = L.op(OP_OR, sc_failed, gv);
          for (size_t k = 0; k < P.aux.size(); ++k) {
            const Vals& old = in[n_in + 2 + k];
            const Vals& nw = r[P.aux[k]];
            aux_new[k].resize(old.size());
            for (size_t e = 0; e < old.size(); ++e) aux_new[k][e] = sel(sc_ls, nw[e], old[e]);
          }
          CNT[0] = SCn[3]; CNT[1] = SCn[4]; CNT[2] = zero; CNT[3] = SCn[5];
        }
        ccu_int o = 0;
        for (casadi_int i = 0; i < n; ++i) lib.builder_output(L.b, o, i, Xn[i]);
        ++o;
        for (casadi_int i = 0; i < n; ++i) lib.builder_output(L.b, o, i, DXn[i]);
        ++o;
        for (casadi_int i = 0; i < NEWTON_SC; ++i) lib.builder_output(L.b, o, i, SCn[i]);
        ++o;
        for (size_t k = 0; k < P.aux.size(); ++k, ++o)
          for (size_t e = 0; e < aux_new[k].size(); ++e) lib.builder_output(L.b, o, static_cast<ccu_int>(e), aux_new[k][e]);
        for (casadi_int i = 0; i < NEWTON_COUNTS; ++i) lib.builder_output(L.b, o, i, CNT[i]);
        P.tape[which] = export_builder(lib, L.b, t_in, t_out);
      } catch (...) {
        lib.builder_destroy(L.b);
        throw;
      }
      lib.builder_destroy(L.b);
    }
    return P;
  }

  int CudaMap::newton_run(const NewtonPlan& P, casadi_int N, const double* const* arg, double* const* res, NewtonBackend& be,
                          casadi_int* n_failed, casadi_int* n_singular, casadi_int launches[2]) {
    const casadi_int n_in = static_cast<casadi_int>(P.nnz_in.size()), n_aux = static_cast<casadi_int>(P.aux.size());
    std::vector<double*> owned;
    auto get = [&](casadi_int count) { double* p = count > 0 ? be.alloc(count) : nullptr; if (p) owned.push_back(p); return p; };
    // inputs other than the unknown: constant during the solve (a null argument reads as zeros in the tape)
    std::vector<const double*> fixed(n_in, nullptr);
    for (casadi_int j = 0; j < n_in; ++j) {
      if (j == P.iin || !arg[j] || P.nnz_in[j] == 0) continue;
      double* d = get(N * P.nnz_in[j]);
      be.upload(d, arg[j], N * P.nnz_in[j]);
      fixed[j] = d;
    }
    // state, double buffered: a launch reads one copy and writes the other
    double *X[2], *DX[2], *SC[2];
    std::vector<double*> AUX[2];
    for (int b = 0; b < 2; ++b) {
      X[b] = get(N * P.n); DX[b] = get(N * P.n); SC[b] = get(N * NEWTON_SC);
      for (casadi_int k = 0; k < n_aux; ++k) AUX[b].push_back(get(N * P.nnz_out[P.aux[k]]));
    }
    be.upload(X[0], arg[P.iin], N * P.n);  // the initial guess (null: zeros)
    be.upload(DX[0], nullptr, N * P.n);
    {
      std::vector<double> sc0(N * NEWTON_SC, 0.);
      for (casadi_int i = 0; i < N; ++i) { sc0[i * NEWTON_SC + 2] = 1.; sc0[i * NEWTON_SC + 3] = 1.; }
      be.upload(SC[0], get_ptr(sc0), N * NEWTON_SC);
    }
    for (casadi_int k = 0; k < n_aux; ++k) if (AUX[0][k]) be.upload(AUX[0][k], nullptr, N * P.nnz_out[P.aux[k]]);
    int cur = 0, flag = 0;
    double counts[NEWTON_COUNTS] = {static_cast<double>(N), 0, 0, 0};
    launches[0] = launches[1] = 0;
    auto launch = [&](int which) {
      std::vector<const double*> a(fixed);
      a[P.iin] = X[cur];
      a.push_back(DX[cur]); a.push_back(SC[cur]);
      for (casadi_int k = 0; k < n_aux; ++k) a.push_back(AUX[cur][k]);
      std::vector<double*> r = {X[1 - cur], DX[1 - cur], SC[1 - cur]};
      for (casadi_int k = 0; k < n_aux; ++k) r.push_back(AUX[1 - cur][k]);
      r.push_back(nullptr);  // the counts: placed by the backend
      int f = be.launch(which, N, a, r, counts);
      cur = 1 - cur;
      launches[which]++;
      return f;
    };
    casadi_int singular = 0;
    for (casadi_int iter = 0; iter < P.max_iter && counts[0] > 0 && !flag; ++iter) {
      flag = launch(0);
      singular += static_cast<casadi_int>(counts[2]);
      while (!flag && counts[1] > 0) flag = launch(1);
    }
    if (!flag) {
      if (res[P.iout]) be.download(res[P.iout], X[cur], N * P.n);
      for (casadi_int k = 0; k < n_aux; ++k)
        if (res[P.aux[k]] && AUX[cur][k]) be.download(res[P.aux[k]], AUX[cur][k], N * P.nnz_out[P.aux[k]]);
    }
    for (double* p : owned) be.release(p);
    // failed: the line search gave up, or still iterating after max_iter iterations (newton.cpp:143-149)
    *n_failed = static_cast<casadi_int>(counts[3] + counts[0]);
    *n_singular = singular;
    if (flag) return flag;
    // (a singular Jacobian is not an error here: Newton::solve ignores the return value of Linsol::nfact, newton.cpp:173-174,
    // and solves with whatever the factorisation holds -- the tapes do the same arithmetic; the count is informational)
    return P.error_on_fail && *n_failed > 0 ? 1 : 0;
  }

  namespace {
    // the state on the device, the tapes evaluated by ccu_map_eval_reduce_device (AoS), the counts through a 4-double buffer
    struct NewtonDevice : public CudaMap::NewtonBackend {
      CudaLib& lib;
      void* tape[2];
      double* d_counts = nullptr;
      std::string error;
      NewtonDevice(CudaLib& l, void* t0, void* t1) : lib(l) { tape[0] = t0; tape[1] = t1; d_counts = alloc(CudaMap::NEWTON_COUNTS); }
      ~NewtonDevice() override { if (d_counts) lib.dev_free(d_counts); }
      double* alloc(casadi_int n) override {
        void* p = lib.dev_malloc(static_cast<ccu_int>(n) * 8);
        casadi_assert(p != nullptr, "Map 'cuda': " + std::string(lib.last_error()));
        return static_cast<double*>(p);
      }
      void release(double* p) override { lib.dev_free(p); }
      void upload(double* dst, const double* src, casadi_int n) override {
        if (n <= 0) return;
        std::vector<double> z;
        if (!src) { z.assign(n, 0.); src = get_ptr(z); }
        casadi_assert(lib.memcpy_h2d(dst, src, static_cast<ccu_int>(n) * 8, nullptr) == 0 && lib.stream_sync(nullptr) == 0,
                      "Map 'cuda': " + std::string(lib.last_error()));
      }
      void download(double* dst, const double* src, casadi_int n) override {
        if (n <= 0) return;
        casadi_assert(lib.memcpy_d2h(dst, src, static_cast<ccu_int>(n) * 8, nullptr) == 0 && lib.stream_sync(nullptr) == 0,
                      "Map 'cuda': " + std::string(lib.last_error()));
      }
      int launch(int which, casadi_int N, const std::vector<const double*>& arg, const std::vector<double*>& res,
                 double* counts) override {
        // outputs: X, DX, SC, aux..., counts -- only the last is summed over the instances
        std::vector<double*> r(res);
        r.back() = d_counts;
        std::vector<int> red(r.size(), 0);
        red.back() = 1;
        if (lib.eval_reduce_device(tape[which], N, get_ptr(arg), get_ptr(r), nullptr, get_ptr(red), 0 /* CCU_LAYOUT_AOS */, nullptr)) return 1;
        if (lib.memcpy_d2h(counts, d_counts, CudaMap::NEWTON_COUNTS * 8, nullptr) || lib.stream_sync(nullptr)) return 1;
        return 0;
      }
    };
  } // namespace

  void CudaMap::init(const Dict& opts) {
    // Map::create passes an empty Dict (map.cpp:43-47): the devices come from the environment.
    //   CASADI_CUDA_DEVICE=k          one device (default 0)
    //   CASADI_CUDA_DEVICES=all|0,1,3 several devices of this process: instances [g*n/G, (g+1)*n/G) on device g,
    //                                 reduce_out sums combined by NCCL inside libcasadi_cuda.so
    if (const char* d = getenv("CASADI_CUDA_DEVICE")) device_ = atoi(d);
    devices_.clear();
    if (const char* d = getenv("CASADI_CUDA_DEVICES")) {
      std::string v = d;
      if (v == "all") {
        CudaLib& l = cuda_lib();
        casadi_assert(l.handle!=nullptr, "Map 'cuda': " + l.error);
        for (int k = 0; k < l.device_count(); ++k) devices_.push_back(k);
      } else {
        size_t pos = 0;
        while (pos <= v.size()) {
          size_t c = v.find(',', pos);
          if (c == std::string::npos) c = v.size();
          if (c > pos) devices_.push_back(atoi(v.substr(pos, c - pos).c_str()));
          pos = c + 1;
        }
      }
    }
    if (devices_.empty()) devices_.push_back(device_);

    // Call the initialization method of the base class (work vectors for one serial evaluation are
    // more than the device path needs: no per-instance host work vector is allocated)
    Map::init(opts);

    export_function();

    // Fail now, not at the first evaluation, if the device library is unusable (there is no fallback)
    CudaLib& lib = cuda_lib();
    casadi_assert(lib.handle!=nullptr, "Map 'cuda': " + lib.error);
  }

  int CudaMap::init_mem(void* mem) const {
    if (Map::init_mem(mem)) return 1;
    auto m = static_cast<CudaMapMemory*>(mem);
    CudaLib& lib = cuda_lib();
    casadi_assert(lib.handle!=nullptr, "Map 'cuda': " + lib.error);
    const Tape& t = tape_;
    std::vector<int> dv = devices_.empty() ? std::vector<int>{device_} : devices_;
    if (newton_) {
      casadi_assert(dv.size() == 1, "Map 'cuda': a mapped Newton rootfinder runs on one device (CASADI_CUDA_DEVICE)");
      for (int k = 0; k < 2; ++k) {
        const Tape& tk = newton_plan_.tape[k];
        m->newton_tape[k] = lib.tape_create(static_cast<ccu_int>(tk.op.size()), get_ptr(tk.op), get_ptr(tk.i0), get_ptr(tk.i1),
                                            get_ptr(tk.i2), get_ptr(tk.d), tk.sz_w, static_cast<ccu_int>(tk.nnz_in.size()),
                                            get_ptr(tk.nnz_in), static_cast<ccu_int>(tk.nnz_out.size()), get_ptr(tk.nnz_out), dv[0]);
        casadi_assert(m->newton_tape[k]!=nullptr, "Map 'cuda': cannot put rootfinder '" + f_.name() + "' on device "
                      + str(dv[0]) + ": " + std::string(lib.last_error()));
      }
      m->add_stat("cuda");
      return 0;
    }
    if (builder_) {
      m->tape = lib.builder_finish_multi(builder_, static_cast<ccu_int>(t.nnz_in.size()), get_ptr(t.nnz_in),
                                         static_cast<ccu_int>(t.nnz_out.size()), get_ptr(t.nnz_out),
                                         static_cast<int>(dv.size()), get_ptr(dv));
    } else {
      m->tape = lib.multi_create(static_cast<ccu_int>(t.op.size()), get_ptr(t.op), get_ptr(t.i0),
                                 get_ptr(t.i1), get_ptr(t.i2), get_ptr(t.d), t.sz_w,
                                 static_cast<ccu_int>(t.nnz_in.size()), get_ptr(t.nnz_in),
                                 static_cast<ccu_int>(t.nnz_out.size()), get_ptr(t.nnz_out),
                                 static_cast<int>(dv.size()), get_ptr(dv));
    }
    casadi_assert(m->tape!=nullptr, "Map 'cuda': cannot put function '" + f_.name() + "' on device(s) "
                  + str(dv) + ": " + std::string(lib.last_error()));
    // FStats (timing.hpp:47-98): the whole call, and the device time of its three phases (they overlap chunk by chunk,
    // so the parts do not add up to the whole) plus the host time spent staging pageable buffers
    m->add_stat("cuda");
    m->add_stat("cuda_h2d");
    m->add_stat("cuda_kernel");
    m->add_stat("cuda_d2h");
    m->add_stat("cuda_stage");
    return 0;
  }

  void CudaMap::free_mem(void *mem) const {
    auto m = static_cast<CudaMapMemory*>(mem);
    if (m->tape) cuda_lib().multi_destroy(m->tape);
    for (int k = 0; k < 2; ++k) if (m->newton_tape[k]) cuda_lib().tape_destroy(m->newton_tape[k]);
    delete m;
  }

  int CudaMap::eval(const double** arg, double** res, casadi_int* iw, double* w, void* mem) const {
    return eval_reduce(arg, res, std::vector<bool>(), std::vector<bool>(), mem);
  }

  int CudaMap::eval_reduce(const double** arg, double** res, const std::vector<bool>& reduce_in,
                           const std::vector<bool>& reduce_out, void* mem) const {
    auto m = static_cast<CudaMapMemory*>(mem);
    CudaLib& lib = cuda_lib();
    m->stats_available = true;  // F.stats() reports the timers of the last call (function_internal.cpp:3168-3175)
    m->fstats.at("cuda").tic();
    if (newton_) {
      casadi_assert(reduce_in.empty() && reduce_out.empty() && in_groups_.empty() && out_groups_.empty(),
                    "Map 'cuda': reductions and derivative layouts are not available for a mapped rootfinder");
      std::vector<int> dv = devices_.empty() ? std::vector<int>{device_} : devices_;
      casadi_assert(lib.set_device(dv[0]) == 0, "Map 'cuda': " + std::string(lib.last_error()));
      casadi_int n_failed = 0, n_singular = 0, launches[2] = {0, 0};
      int flag;
      {
        NewtonDevice be(lib, m->newton_tape[0], m->newton_tape[1]);
        flag = newton_run(newton_plan_, n_*rep_, arg, res, be, &n_failed, &n_singular, launches);
      }
      m->fstats.at("cuda").toc();
      // Rootfinder::eval raises when an instance failed and error_on_fail is set (rootfinder.cpp:294-296).  A numerically
      // singular Jacobian is not an error by itself: Newton::solve ignores the failure of Linsol::nfact (newton.cpp:173-174)
      // and the step is whatever casadi_qr_solve makes of the factorisation -- reproduced bit for bit by the tapes
      if (n_singular > 0 && verbose_)
        casadi_message("Map 'cuda': the Jacobian of rootfinder '" + leaf_.name() + "' was numerically singular in "
                       + str(n_singular) + " factorisation(s)");
      if (n_failed > 0 && newton_plan_.error_on_fail)
        casadi_error("rootfinder process failed for " + str(n_failed) + " of " + str(n_*rep_) + " instances of '" + leaf_.name()
                     + "'. Set 'error_on_fail' option to false to ignore this error.");
      if (flag) casadi_warning("Map 'cuda' evaluation of '" + f_.name() + "' failed: " + std::string(lib.last_error()));
      return flag;
    }
    // Same contract as Map::eval_gen (map.cpp:141-157): instance i of input j is arg[j]+i*nnz_in(j);
    // null arg[j] reads as zero, null res[j] is not computed.  With reductions (MapSum::eval_gen, mapsum.cpp:154-186):
    // a reduced input is ONE instance read by all, a reduced output is the sum over the instances.
    int flag;
    double n_failed = 0;
    const bool reduced = !reduce_in.empty() || !reduce_out.empty();
    std::vector<int> gi(in_groups_.begin(), in_groups_.end()), go(out_groups_.begin(), out_groups_.end());
    const int* gip = gi.empty() ? nullptr : get_ptr(gi);
    const int* gop = go.empty() ? nullptr : get_ptr(go);
    // reductions are defined over whole instances of f_ (CudaMapSum builds its map with keep_nested())
    casadi_assert(!reduced || rep_ == 1, "Map 'cuda': reductions over a flattened nested map");
    if (has_flag_ || reduced) {
      // the failure count of lowered linear solvers is one extra, summed output
      std::vector<double*> r(res, res + n_out_);
      std::vector<int> red_out(n_out_, 0), red_in(n_in_, 0);
      for (size_t j=0; j<reduce_out.size() && j<red_out.size(); ++j) red_out[j] = reduce_out[j];
      for (size_t j=0; j<reduce_in.size() && j<red_in.size(); ++j) red_in[j] = reduce_in[j];
      if (has_flag_) {
        r.push_back(&n_failed);
        red_out.push_back(1);
        if (gop) { go.push_back(1); gop = get_ptr(go); }
      }
      flag = lib.multi_eval_host(m->tape, n_*rep_, arg, get_ptr(r), reduce_in.empty() ? nullptr : get_ptr(red_in),
                                 get_ptr(red_out), gip, gop);
    } else {
      flag = lib.multi_eval_host(m->tape, n_*rep_, arg, res, nullptr, nullptr, gip, gop);
    }
    m->fstats.at("cuda").toc();
    {
      // with several devices: the slowest device of each phase
      double worst[4] = {0, 0, 0, 0};
      for (int g = 0; g < lib.multi_size(m->tape); ++g) {
        double st[6] = {0, 0, 0, 0, 0, 0};
        if (lib.last_eval_stats(lib.multi_tape(m->tape, g), st) == 0)
          for (int k = 0; k < 4; ++k) worst[k] = std::max(worst[k], st[k]);
      }
      const char* names[4] = {"cuda_h2d", "cuda_kernel", "cuda_d2h", "cuda_stage"};
      for (int k = 0; k < 4; ++k) {
        FStats& fs = m->fstats.at(names[k]);
        fs.n_call += 1;
        fs.t_wall += 1e-3 * worst[k];
      }
    }
    if (!flag && n_failed > 0) {
      casadi_warning("Map 'cuda': linear solver factorization failed for " + str(static_cast<casadi_int>(n_failed))
                     + " of " + str(n_) + " instances of '" + f_.name() + "'");
      return 1;
    }
    if (flag) {
      casadi_warning("Map 'cuda' evaluation of '" + f_.name() + "' failed: " + std::string(lib.last_error()));
      return 1;
    }
    return 0;
  }

} // namespace casadi
