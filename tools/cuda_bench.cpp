// End-to-end benchmark of the PLUGIN (MEASUREMENT INFRASTRUCTURE): `f.map(N, "cuda")` created and evaluated through the
// reference's own public C++ API -- Function::map -> Map::create -> CudaMap (casadi_b200/host/cuda_map.cpp), then
// F(arg, res, iw, w, 0) -> FunctionInternal::eval_gen -> CudaMap::eval -- inside the relinked reference library
// (tests/integration/build_integration.py), with the buffers a CasADi caller owns: ordinary pageable std::vector
// storage (default), the same storage page-locked in place by the library ("registered", CCU_HOST_REGISTER=1) or pinned
// memory ("pinned").  Host<->device copies are inside the timed region; construction
// (Map sparsity repmat + tape export + specialisation) is timed separately.
//
// usage: cuda_bench <workload> <n> <reps> <warmup> [pageable|pinned|registered] [reduce]
//   workload: cartpole | quad | quad_jac | quad_ms (= quad then quad_jac) | rocket_hess | mc | kkt_ldl | kkt_qr
//   reduce:   map with every output summed over the instances (Function::map(name, "cuda", n, {}, all outputs)):
//             BASELINE config 4 (mapaccum rollout with reduce_out)
// prints one JSON line.
#include <dlfcn.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bench_models.hpp"

using namespace casadi;

namespace {

struct HostAlloc {
  bool pinned = false;
  void* (*malloc_host)(long long) = nullptr;
  int (*free_host)(void*) = nullptr;
  std::vector<void*> owned;
  double* get(size_t n) {
    n = std::max<size_t>(n, 1);
    void* p = pinned ? malloc_host(static_cast<long long>(n * sizeof(double))) : std::malloc(n * sizeof(double));
    casadi_assert(p != nullptr, "host allocation of " + str(n * 8) + " bytes failed");
    std::memset(p, 0, n * sizeof(double));  // touch every page: the timed region must not include first-touch faults
    owned.push_back(p);
    return static_cast<double*>(p);
  }
  ~HostAlloc() { for (void* p : owned) { if (pinned) free_host(p); else std::free(p); } }
};

struct Job {
  Function f, F;  // the function and its device map
  bool reduce = false;
  long long n = 0;
  std::vector<double*> in, out;
  std::vector<const double*> arg;
  std::vector<double*> res;
  std::vector<casadi_int> iw;
  std::vector<double> w;
  double construct_s = 0;
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: cuda_bench <workload> <n> <reps> <warmup> [pageable|pinned|registered] [reduce]\n");
    return 2;
  }
  const std::string wl = argv[1];
  const long long n = atoll(argv[2]);
  const int reps = atoi(argv[3]), warm = atoi(argv[4]);
  const std::string memkind = argc > 5 ? argv[5] : "pageable";
  const bool reduce = argc > 6 && std::string(argv[6]) == "reduce";
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
    HostAlloc A;
    A.pinned = memkind == "pinned";
    casadi_assert(memkind == "pageable" || memkind == "pinned" || memkind == "registered", "buffers: pageable | pinned | registered");
    // "registered": the same malloc buffers, page-locked in place by the library the first time it sees them (the warm-up
    // call), include/casadi_cuda.h: ccu_host_register.  The maps are destroyed (and un-register) before A frees the buffers.
    if (memkind == "registered") setenv("CCU_HOST_REGISTER", "1", 1);
    if (A.pinned) {
      const char* lib = getenv("CASADI_CUDA_LIB");
      void* h = dlopen(lib ? lib : "libcasadi_cuda.so", RTLD_NOW | RTLD_GLOBAL);
      casadi_assert(h != nullptr, "cannot load the device library for pinned allocations");
      A.malloc_host = reinterpret_cast<void* (*)(long long)>(dlsym(h, "ccu_malloc_host"));
      A.free_host = reinterpret_cast<int (*)(void*)>(dlsym(h, "ccu_free_host"));
      casadi_assert(A.malloc_host && A.free_host, "ccu_malloc_host missing");
    }
    std::string kind;
    std::vector<Function> fs = ccu_models::bench_workload(wl, &kind);
    const unsigned long long seed = kind == "cartpole" ? 1 : kind == "quad" ? 2 : kind == "rocket" ? 3 : kind == "mc" ? 4 : 5;
    std::vector<Job> jobs(fs.size());
    long long h2d = 0, d2h = 0;
    for (size_t q = 0; q < fs.size(); ++q) {
      Job& j = jobs[q];
      j.f = fs[q];
      j.n = n;
      j.reduce = reduce;
      const double t0 = now();
      if (reduce) {
        std::vector<casadi_int> rout;
        for (casadi_int k = 0; k < j.f.n_out(); ++k) rout.push_back(k);
        j.F = j.f.map("bench_" + j.f.name(), "cuda", n, std::vector<casadi_int>(), rout);
      } else {
        j.F = j.f.map(n, "cuda");
      }
      j.construct_s = now() - t0;
      std::vector<std::vector<double>> in;
      ccu_models::bench_inputs(j.f, n, seed, kind, in);
      for (casadi_int k = 0; k < j.f.n_in(); ++k) {
        j.in.push_back(A.get(in[k].size()));
        std::memcpy(j.in.back(), in[k].data(), in[k].size() * sizeof(double));
        h2d += static_cast<long long>(in[k].size()) * 8;
      }
      for (casadi_int k = 0; k < j.f.n_out(); ++k) {
        const size_t cnt = static_cast<size_t>(j.f.nnz_out(k)) * (reduce ? 1 : n);
        j.out.push_back(A.get(cnt));
        d2h += static_cast<long long>(cnt) * 8;
      }
      j.arg.assign(j.F.sz_arg(), nullptr);
      j.res.assign(j.F.sz_res(), nullptr);
      j.iw.resize(j.F.sz_iw());
      j.w.resize(j.F.sz_w());
      for (casadi_int k = 0; k < j.f.n_in(); ++k) j.arg[k] = j.in[k];
      for (casadi_int k = 0; k < j.f.n_out(); ++k) j.res[k] = j.out[k];
    }
    std::vector<double> secs;
    for (int r = 0; r < reps + warm; ++r) {
      const double t0 = now();
      for (Job& j : jobs) {
        int flag = j.F(j.arg.data(), j.res.data(), j.iw.data(), j.w.data(), 0);
        casadi_assert(flag == 0, "evaluation of " + j.F.name() + " failed");
      }
      const double dt = now() - t0;
      if (r >= warm) secs.push_back(dt);
    }
    // optional sweep of the host pipeline's chunk size on the same maps (CUDA_BENCH_CHUNK_SWEEP="65536,131072,..."):
    // the library reads CCU_HOST_CHUNK at every call, so no map has to be rebuilt
    if (const char* sweep = getenv("CUDA_BENCH_CHUNK_SWEEP")) {
      std::string list = sweep;
      size_t pos = 0;
      while (pos < list.size()) {
        size_t c = list.find(',', pos);
        if (c == std::string::npos) c = list.size();
        const std::string item = list.substr(pos, c - pos);
        pos = c + 1;
        if (item.empty()) continue;
        setenv("CCU_HOST_CHUNK", item.c_str(), 1);
        std::vector<double> ss;
        for (int r = 0; r < reps + 1; ++r) {
          const double t0 = now();
          for (Job& j : jobs) casadi_assert(j.F(j.arg.data(), j.res.data(), j.iw.data(), j.w.data(), 0) == 0, "evaluation failed");
          if (r >= 1) ss.push_back(now() - t0);
        }
        std::sort(ss.begin(), ss.end());
        fprintf(stderr, "chunk_sweep %s %s: chunk %s -> %.6f s, %.4g evals/s\n", wl.c_str(), memkind.c_str(), item.c_str(),
                ss[ss.size() / 2], static_cast<double>(n) / ss[ss.size() / 2]);
      }
      unsetenv("CCU_HOST_CHUNK");
    }
    // parity of the timed run's results: first / last instances against the reference's serial evaluation of f
    double worst = 0;
    for (Job& j : jobs) {
      if (j.reduce) continue;
      for (long long i : {0ll, 1ll, n / 2, n - 1}) {
        if (i < 0 || i >= n) continue;
        std::vector<const double*> a(j.f.sz_arg(), nullptr);
        std::vector<double*> r(j.f.sz_res(), nullptr);
        std::vector<casadi_int> iw(j.f.sz_iw());
        std::vector<double> w(j.f.sz_w());
        std::vector<std::vector<double>> o(j.f.n_out());
        for (casadi_int k = 0; k < j.f.n_in(); ++k) a[k] = j.in[k] + i * j.f.nnz_in(k);
        for (casadi_int k = 0; k < j.f.n_out(); ++k) { o[k].resize(j.f.nnz_out(k)); r[k] = o[k].data(); }
        j.f(a.data(), r.data(), iw.data(), w.data(), 0);
        for (casadi_int k = 0; k < j.f.n_out(); ++k)
          for (casadi_int e = 0; e < j.f.nnz_out(k); ++e) {
            const double got = j.out[k][i * j.f.nnz_out(k) + e], want = o[k][e];
            const double err = std::fabs(got - want) / std::max(std::fabs(want), 1.0);
            worst = std::max(worst, (got == want || (got != got && want != want)) ? 0.0 : err);
          }
      }
    }
    double total = 0, construct = 0;
    for (double v : secs) total += v;
    std::sort(secs.begin(), secs.end());
    for (Job& j : jobs) construct += j.construct_s;
    // FStats of the device maps (function_internal.cpp:986-1011): h2d / kernel / d2h split recorded by CudaMap
    std::string stats = "{";
    for (Job& j : jobs) {
      Function cm = j.F;
      if (!cm.is_a("CudaMap", true))
        for (const std::string& nm : j.F.get_function()) if (j.F.get_function(nm).is_a("CudaMap", true) || j.F.get_function(nm).class_name() == "CudaMapSum") cm = j.F.get_function(nm);
      Dict st;
      try { st = cm.stats(); } catch (std::exception&) { continue; }  // (a map called from an MX wrapper used another memory object)
      for (auto&& e : st) {
        if (e.first.rfind("t_wall_", 0) != 0) continue;
        if (stats.size() > 1) stats += ", ";
        stats += "\"" + j.f.name() + "." + e.first + "\": " + str(e.second.to_double());
      }
    }
    stats += "}";
    int registered = 0;  // buffers the library holds page-locked ("registered" mode; 0 when the registration was refused)
    {
      const char* lib = getenv("CASADI_CUDA_LIB");
      void* h = dlopen(lib ? lib : "libcasadi_cuda.so", RTLD_NOW | RTLD_GLOBAL);
      if (h) if (auto cnt = reinterpret_cast<int (*)()>(dlsym(h, "ccu_host_registered_count"))) registered = cnt();
    }
    printf("{\"workload\": \"%s\", \"n\": %lld, \"memory\": \"%s\", \"reduce\": %s, \"reps\": %d, \"secs_median\": %.6f, "
           "\"secs_best\": %.6f, \"secs_total\": %.6f, \"evals_per_s\": %.6g, \"construct_s\": %.3f, "
           "\"h2d_bytes_per_step\": %lld, \"d2h_bytes_per_step\": %lld, \"parity_rel_err\": %.3g, \"registered_buffers\": %d, \"fstats\": %s}\n",
           wl.c_str(), n, memkind.c_str(), reduce ? "true" : "false", reps, secs[secs.size() / 2], secs.front(), total,
           static_cast<double>(n) * reps / total, construct, h2d, d2h, worst, registered, stats.c_str());
    fflush(stdout);
  } catch (std::exception& e) {
    fprintf(stderr, "cuda_bench: %s\n", e.what());
    return 1;
  }
  return 0;
}
