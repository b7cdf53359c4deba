/*
 * CudaMapSum -- `MapSum::create(name, "cuda", f, n, reduce_in, reduce_out)`: the "cuda" parallelization of MapSum
 * (casadi/core/mapsum.hpp:40-232), which the reference only offers as "serial" (mapsum.cpp:41-54).
 *
 * NEW file for casadi/core/ (not a copy of reference code); the reference-side change is in
 * casadi_map_cuda.patch.  The device work is CudaMap's: CudaMapSum owns a CudaMap of the same function and calls
 * its reducing entry point, so tape export, MX lowering, memory objects and the no-CPU-fallback rule are shared.
 * Sums are evaluated by the fixed-shape tree of libcasadi_cuda.so (1024-instance blocks, then a binary tree):
 * the same value as MapSum::eval_gen's sequential sum (mapsum.cpp:170-184) up to floating-point summation order.
 */
#ifndef CASADI_CUDA_MAPSUM_HPP
#define CASADI_CUDA_MAPSUM_HPP

#include "mapsum.hpp"

/// \cond INTERNAL

namespace casadi {

  class CASADI_EXPORT CudaMapSum : public MapSum {
    friend class MapSum;
  public:
    // Constructor (use MapSum::create(name, "cuda", ...))
    CudaMapSum(const std::string& name, const Function& f, casadi_int n,
               const std::vector<bool>& reduce_in, const std::vector<bool>& reduce_out);

    ~CudaMapSum() override;

    std::string class_name() const override {return "CudaMapSum";}

    bool is_a(const std::string& type, bool recursive) const override;

    /// Keeps MapSum::get_forward/get_reverse (mapsum.cpp:304,367) on the GPU
    std::string parallelization() const override { return "cuda"; }

    void init(const Dict& opts) override;

    int eval(const double** arg, double** res, casadi_int* iw, double* w, void* mem) const override;

    bool has_codegen() const override { return false;}

  protected:
    explicit CudaMapSum(DeserializingStream& s);

  private:
    // f_.map(n_, "cuda"): owns the device program
    Function map_;
  };

} // namespace casadi
/// \endcond

#endif // CASADI_CUDA_MAPSUM_HPP
