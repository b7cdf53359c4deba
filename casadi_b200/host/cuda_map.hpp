/*
 * CudaMap -- `Function::map(n, "cuda")`: a fourth Map parallelization next to "serial", "openmp" and
 * "thread" (casadi/core/map.hpp:39-330).
 *
 * This file and cuda_map.cpp are NEW files for casadi/core/ (they are not copies of reference code); the
 * reference-side change is casadi_map_cuda.patch (a "cuda" branch in Map::create, a "CudaMap" branch in
 * Map::deserialize, two CMake lines).  See INTEGRATION.md.
 *
 * CudaMap exports the SXFunction instruction tape of the mapped function through the public accessors
 * (function.hpp:1114-1138), hands it to libcasadi_cuda.so through the C ABI of include/casadi_cuda.h
 * (dlopen'ed on first use, so libcasadi itself gains no CUDA dependency) and evaluates all n instances
 * on the GPU.  There is no CPU fallback: functions that cannot run on the device fail at init.
 */
#ifndef CASADI_CUDA_MAP_HPP
#define CASADI_CUDA_MAP_HPP

#include "map.hpp"

/// \cond INTERNAL

namespace casadi {

  /** Memory object of a CudaMap: one compiled tape (device program, staging buffers, streams) per
      checked-out memory, so concurrent evaluations on distinct memory objects do not share state
      (function_internal.hpp:184-194). */
  struct CASADI_EXPORT CudaMapMemory : public FunctionMemory {
    void* tape;  // ccu_multi*: the compiled tape on every device of the map
    void* newton_tape[2];  // ccu_tape*: direction and line-search tapes of a mapped Newton rootfinder
    CudaMapMemory() : tape(nullptr) { newton_tape[0] = newton_tape[1] = nullptr; }
  };

  class CASADI_EXPORT CudaMap : public Map {
    friend class Map;
  public:
    // Constructor (use Map::create("cuda", f, n))
    CudaMap(const std::string& name, const Function& f, casadi_int n);

    ~CudaMap() override;

    std::string class_name() const override {return "CudaMap";}

    bool is_a(const std::string& type, bool recursive) const override;

    /// Type of parallellization: keeps Map::get_forward/get_reverse (map.cpp:226,280) on the GPU
    std::string parallelization() const override { return "cuda"; }

    void init(const Dict& opts) override;

    /// Evaluate the function numerically: all n instances on the device
    int eval(const double** arg, double** res, casadi_int* iw, double* w, void* mem) const override;

    /** The same evaluation with MapSum semantics (mapsum.cpp:154-186): reduce_in[j] -- arg[j] is one instance read by
        every evaluation; reduce_out[j] -- res[j] receives the sum over the instances (fixed-shape tree on the device).
        Empty vectors = plain map.  Used by CudaMapSum. */
    int eval_reduce(const double** arg, double** res, const std::vector<bool>& reduce_in,
                    const std::vector<bool>& reduce_out, void* mem) const;

    /** Derivatives of the map: df.map(n, "cuda") whose seed / sensitivity blocks are read and written in the
        direction-major layout of a derivative function directly (ccu_multi_eval_host_grouped), instead of the
        reference's GetNonzeros column permutations on the host around the mapped call (map.cpp:219-325) */
    Function get_forward(casadi_int nfwd, const std::string& name, const std::vector<std::string>& inames,
                         const std::vector<std::string>& onames, const Dict& opts) const override;
    Function get_reverse(casadi_int nadj, const std::string& name, const std::vector<std::string>& inames,
                         const std::vector<std::string>& onames, const Dict& opts) const override;

    /** Piece-major caller layout: input / output j of an instance is groups[j] pieces, piece d of instance k at
        (d*n + k) * nnz/groups[j] (how a derivative function lays out its nfwd / nadj blocks).  Call before init. */
    void set_groups(const std::vector<casadi_int>& in_groups, const std::vector<casadi_int>& out_groups) {
      in_groups_ = in_groups; out_groups_ = out_groups;
    }

    void serialize_body(SerializingStream &s) const override;

    /// No C code generation for the device path
    bool has_codegen() const override { return false;}

    void* alloc_mem() const override { return new CudaMapMemory(); }
    int init_mem(void* mem) const override;
    void free_mem(void *mem) const override;

    /** The exported tape of the mapped function (reference layout, sx_function.hpp:37-44) */
    struct Tape {
      std::vector<int> op, i0, i1, i2;
      std::vector<double> d;
      casadi_int sz_w;
      std::vector<casadi_int> nnz_in, nnz_out;
    };
    static Tape export_tape(const Function& f);

    /** The tape one device thread evaluates for `f`: export_tape for an SX function or an MX function that expands;
        otherwise the function lowered node by node through the tape builder (Linsol calls, fixed-step integrators).
        Needs libcasadi_cuda.so but no device: host-side checks evaluate it instance by instance.  A last extra output
        of one nonzero is present when the lowering counts failed QR factorisations. */
    static Tape lowered_tape(const Function& f);

    /** The Newton rootfinder under the map (casadi/solvers/newton.cpp:130-246).  Newton::solve has a data
        dependent iteration count and line search per instance, so it is not one straight-line tape: it is two --
        tape 0, the Newton direction (jac_g_x, convergence test on max|F|, Linsol factorise + solve, test on the step),
        and tape 1, one line-search trial (x - alpha*dx, g, acceptance test, alpha halved) -- each evaluated for ALL
        instances per launch with the instance's progress flags as data: an instance that has finished, or is not in
        the phase the launch belongs to, keeps every value (bit-exact select).  The host only counts: after every
        launch it reads the sums of the flags (a reduce_out output) and decides whether another line-search trial or
        another Newton iteration is needed, up to max_iter.  All arithmetic of the solver runs on the device in the
        reference's order, so every instance gets the bits Newton::solve gives it.
        State of an instance (AoS, one array each): the iterate X (stored in the rootfinder input iin), the step DX,
        the scalars SC = (abstol, abstolStep, alpha, active, in_line_search, failed), the auxiliary outputs. */
    struct NewtonPlan {
      Tape tape[2];          // inputs: the rootfinder's inputs (iin carries X), DX, SC, aux...; outputs: X, DX, SC, aux..., counts
      casadi_int n = 0, iin = 0, iout = 0, max_iter = 0;
      bool line_search = true, error_on_fail = true;
      std::vector<casadi_int> nnz_in, nnz_out;  // of the rootfinder
      std::vector<casadi_int> aux;              // the rootfinder outputs other than iout, in order
      std::string name;
    };
    enum { NEWTON_SC = 6, NEWTON_COUNTS = 4 };  // counts: active, in line search, singular Jacobians, failed
    /** Where the state lives and how a tape is evaluated over it: the device (CudaMap::eval) or, in the host-side
        checks, plain arrays evaluated by the oracle. */
    struct NewtonBackend {
      virtual ~NewtonBackend() {}
      virtual double* alloc(casadi_int n_doubles) = 0;
      virtual void release(double* p) = 0;
      virtual void upload(double* dst, const double* src, casadi_int n) = 0;  // src == nullptr: zeros
      virtual void download(double* dst, const double* src, casadi_int n) = 0;
      /// evaluate tape `which` over N instances (AoS arrays; a null array reads as zeros / is not written); the last
      /// output -- res.back(), null on entry -- is the counts: summed over the instances into counts[NEWTON_COUNTS]
      virtual int launch(int which, casadi_int N, const std::vector<const double*>& arg, const std::vector<double*>& res,
                         double* counts) = 0;
    };
    static bool is_newton(const Function& f);
    static NewtonPlan newton_plan(const Function& rootfinder);
    /** Newton::solve for N instances; arg / res as Map::eval_gen.  Returns 0, or 1 when an instance failed and the
        rootfinder has error_on_fail (Rootfinder::eval raises in that case, rootfinder.cpp:294-296); n_failed and the
        launch counts are reported. */
    static int newton_run(const NewtonPlan& P, casadi_int N, const double* const* arg, double* const* res, NewtonBackend& be,
                          casadi_int* n_failed, casadi_int* n_singular, casadi_int launches[2]);

    /** Keep a mapped function that is itself a Map as ONE instance (its inner map is expanded into the tape) instead
        of flattening it to n*d device instances.  CudaMapSum needs this: a reduced input belongs to a whole
        instance of f_ and a reduced output is the sum of whole f_ outputs (mapsum.cpp:154-186).  Call before init. */
    void keep_nested() { flatten_ = false; }

  protected:
    explicit CudaMap(DeserializingStream& s);

  private:
    // The SX function whose tape runs on the device (f_ itself, or f_.expand() for an MX function)
    Function sx_;
    /** The function one device thread evaluates and how many of its instances one instance of f_ holds: nested
        maps (Function::map with max_num_threads, function.cpp:829-858) are flattened */
    Function leaf_;
    casadi_int rep_;
    bool flatten_;
    Tape tape_;
    int device_;
    std::vector<casadi_int> in_groups_, out_groups_;  // piece-major layout of derivative blocks (empty = none)
    std::vector<int> devices_;  // CASADI_CUDA_DEVICES: the devices this map is sharded over (default: device_ alone)
    // MX functions that cannot be expanded (e.g. Linsol calls) are lowered node by node through the tape
    // builder of libcasadi_cuda.so; the recorded program lives in builder_ (one compiled tape per memory)
    void* builder_;
    bool has_flag_;  // extra summed output: instances whose linear solver factorization failed
    bool newton_;    // the leaf is a Newton rootfinder: evaluated by newton_run on device-resident state
    NewtonPlan newton_plan_;
    void export_function();
    void lower_mx();
  };

} // namespace casadi
/// \endcond

#endif // CASADI_CUDA_MAP_HPP
