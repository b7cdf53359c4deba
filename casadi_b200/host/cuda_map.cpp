/*
 * CudaMap -- see cuda_map.hpp.  New file for casadi/core/.
 */
#include "cuda_map.hpp"

#include <dlfcn.h>

#include <cstdlib>
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
      void* (*tape_create)(ccu_int, const int*, const int*, const int*, const int*, const double*, ccu_int,
                           ccu_int, const ccu_int*, ccu_int, const ccu_int*, int) = nullptr;
      void (*tape_destroy)(void*) = nullptr;
      int (*map_eval_host)(void*, ccu_int, const double* const*, double* const*) = nullptr;
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
        lib.tape_create = reinterpret_cast<decltype(lib.tape_create)>(sym("ccu_tape_create"));
        lib.tape_destroy = reinterpret_cast<decltype(lib.tape_destroy)>(sym("ccu_tape_destroy"));
        lib.map_eval_host = reinterpret_cast<decltype(lib.map_eval_host)>(sym("ccu_map_eval_host"));
        if (!ok) {
          lib.error = "'" + name + "' does not export the casadi_cuda.h entry points";
          lib.handle = nullptr;
        } else if (lib.abi_version() != 2) {
          lib.error = "'" + name + "' has ABI version " + str(lib.abi_version()) + ", expected 2";
          lib.handle = nullptr;
        }
      });
      return lib;
    }
  } // namespace

  CudaMap::CudaMap(const std::string& name, const Function& f, casadi_int n)
    : Map(name, f, n), device_(0) {
  }

  CudaMap::CudaMap(DeserializingStream& s) : Map(s), device_(0) {
    // The device program is not serialized (Map::serialize_body packs f_ and n_ only, map.cpp:94-98):
    // it is re-exported from f_, exactly like a freshly created map
    export_function();
  }

  CudaMap::~CudaMap() {
    clear_mem();
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

  void CudaMap::export_function() {
    if (f_.is_a("SXFunction")) {
      sx_ = f_;
    } else {
      // An MX function whose nodes all have an SX evaluation (mapaccum/fold towers, wrapped maps)
      // collapses to one SX tape; anything else (e.g. a Linsol call, solve_impl.hpp:57-73) cannot
      try {
        sx_ = f_.expand();
      } catch (std::exception& e) {
        casadi_error("Map 'cuda': function '" + f_.name() + "' (" + f_.class_name() + ") is not an SX "
                     "function and cannot be expanded into one: " + std::string(e.what()));
      }
    }
    casadi_assert(!sx_.has_free(), "Map 'cuda': function '" + f_.name() + "' has free variables "
                  + str(sx_.get_free()) + " and cannot be evaluated");
    tape_ = export_tape(sx_);
  }

  void CudaMap::init(const Dict& opts) {
    // Map::create passes an empty Dict (map.cpp:43-47): the device comes from the environment
    if (const char* d = getenv("CASADI_CUDA_DEVICE")) device_ = atoi(d);

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
    m->tape = lib.tape_create(static_cast<ccu_int>(t.op.size()), get_ptr(t.op), get_ptr(t.i0),
                              get_ptr(t.i1), get_ptr(t.i2), get_ptr(t.d), t.sz_w,
                              static_cast<ccu_int>(t.nnz_in.size()), get_ptr(t.nnz_in),
                              static_cast<ccu_int>(t.nnz_out.size()), get_ptr(t.nnz_out), device_);
    casadi_assert(m->tape!=nullptr, "Map 'cuda': cannot put function '" + f_.name() + "' on device "
                  + str(device_) + ": " + std::string(lib.last_error()));
    m->add_stat("cuda");
    return 0;
  }

  void CudaMap::free_mem(void *mem) const {
    auto m = static_cast<CudaMapMemory*>(mem);
    if (m->tape) cuda_lib().tape_destroy(m->tape);
    delete m;
  }

  int CudaMap::eval(const double** arg, double** res, casadi_int* iw, double* w, void* mem) const {
    auto m = static_cast<CudaMapMemory*>(mem);
    CudaLib& lib = cuda_lib();
    m->fstats.at("cuda").tic();
    // Same contract as Map::eval_gen (map.cpp:141-157): instance i of input j is arg[j]+i*nnz_in(j);
    // null arg[j] reads as zero, null res[j] is not computed
    int flag = lib.map_eval_host(m->tape, n_, arg, res);
    m->fstats.at("cuda").toc();
    if (flag) {
      casadi_warning("Map 'cuda' evaluation of '" + f_.name() + "' failed: " + std::string(lib.last_error()));
      return 1;
    }
    return 0;
  }

} // namespace casadi
