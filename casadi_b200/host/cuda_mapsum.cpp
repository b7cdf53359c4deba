/*
 * CudaMapSum -- see cuda_mapsum.hpp.  New file for casadi/core/.
 */
#include "cuda_mapsum.hpp"
#include "cuda_map.hpp"

namespace casadi {

  CudaMapSum::CudaMapSum(const std::string& name, const Function& f, casadi_int n,
                         const std::vector<bool>& reduce_in, const std::vector<bool>& reduce_out)
    : MapSum(name, f, n, reduce_in, reduce_out) {
  }

  // f_.map(n_, "cuda") without the flattening of nested maps: the reductions act on whole instances of f_
  static Function make_map(const Function& f, casadi_int n) {
    CudaMap* cm = new CudaMap("cudamap" + str(n) + "_" + f.name(), f, n);
    cm->keep_nested();
    return Function::create(cm, Dict());
  }

  CudaMapSum::CudaMapSum(DeserializingStream& s) : MapSum(s) {
    // the device program is not serialized: it is rebuilt from f_ like for a freshly created object
    map_ = make_map(f_, n_);
  }

  CudaMapSum::~CudaMapSum() {
    clear_mem();
  }

  bool CudaMapSum::is_a(const std::string& type, bool recursive) const {
    return type=="CudaMapSum" || (recursive && MapSum::is_a(type, recursive));
  }

  void CudaMapSum::init(const Dict& opts) {
    MapSum::init(opts);
    // Raises here when the function cannot run on the device or no device is present (no CPU fallback)
    map_ = make_map(f_, n_);
  }

  int CudaMapSum::eval(const double** arg, double** res, casadi_int* iw, double* w, void* mem) const {
    const CudaMap* cm = map_.get<CudaMap>();
    casadi_assert(cm!=nullptr, "CudaMapSum: internal map is not a CudaMap");
    casadi_int m = map_.checkout();
    int flag;
    try {
      flag = cm->eval_reduce(arg, res, reduce_in_, reduce_out_, map_.memory(static_cast<int>(m)));
    } catch (...) {
      map_.release(static_cast<int>(m));
      throw;
    }
    map_.release(static_cast<int>(m));
    return flag;
  }

} // namespace casadi
