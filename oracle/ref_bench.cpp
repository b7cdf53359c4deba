// Reference-arm timing tool (TEST / MEASUREMENT INFRASTRUCTURE; not on the product path).
//
// Links against the UNMODIFIED reference built by oracle/build_ref.py and times the reference's own
// CPU implementation of the hot path on this host: Map::eval ("serial", casadi/core/map.cpp:327-334)
// and OmpMap::eval ("openmp", map.cpp:340-386) over the BASELINE models of oracle/models.hpp, through
// the buffer API Function::operator()(arg,res,iw,w,mem) (function.cpp:1708-1738), as BASELINE.md 3
// prescribes.  OpenMP runs as f.map(n/T,"serial").map(T,"openmp") -- what
// Function::map(n,par,max_num_threads) builds (function.cpp:829-858) -- because plain
// f.map(n,"openmp") allocates sz_w*n work doubles and n memory objects (map.cpp:352-353,439-442).
//
// usage: ref_bench <workload> <n> <serial|openmp> <threads> <reps> [warmup=1]
//   workload: cartpole | quad | quad_jac | quad_ms (= quad then quad_jac) | rocket_hess | mc | kkt_ldl | kkt_qr
// prints one JSON line.
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "models.hpp"

using namespace casadi;

struct Job {
  Function F;  // the mapped function
  std::vector<std::vector<double>> in, out;
  std::vector<const double*> arg;
  std::vector<double*> res;
  std::vector<casadi_int> iw;
  std::vector<double> w;
};

static Job make_job(const Function& f, long long n, const std::string& mode, int T, unsigned long long seed,
                    const std::string& kind) {
  Job j;
  if (mode == "openmp") {
    casadi_assert(n % T == 0, "n must be a multiple of the thread count");
    j.F = f.map(n / T, "serial").map(T, "openmp");
  } else {
    j.F = f.map(n, "serial");
  }
  ccu_models::bench_inputs(f, n, seed, kind, j.in);
  j.out.resize(f.n_out());
  j.arg.assign(j.F.sz_arg(), nullptr);
  j.res.assign(j.F.sz_res(), nullptr);
  j.iw.resize(j.F.sz_iw());
  j.w.resize(j.F.sz_w());
  for (casadi_int k = 0; k < f.n_in(); ++k) j.arg[k] = j.in[k].data();
  for (casadi_int k = 0; k < f.n_out(); ++k) {
    j.out[k].assign(n * f.nnz_out(k), 0.0);
    j.res[k] = j.out[k].data();
  }
  return j;
}

static void run(Job& j) {
  // (re)bind the buffers: a Job may have been copied when the job list grew
  for (size_t k = 0; k < j.in.size(); ++k) j.arg[k] = j.in[k].data();
  for (size_t k = 0; k < j.out.size(); ++k) j.res[k] = j.out[k].data();
  int flag = j.F(j.arg.data(), j.res.data(), j.iw.data(), j.w.data(), 0);
  casadi_assert(flag == 0, "reference evaluation failed");
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: ref_bench <workload> <n> <serial|openmp> <threads> <reps> [warmup=1]\n");
    return 2;
  }
  const std::string wl = argv[1], mode = argv[3];
  long long n = atoll(argv[2]);
  int T = atoi(argv[4]);
  const int reps = atoi(argv[5]);
  const int warm = argc > 6 ? atoi(argv[6]) : 1;
  if (T <= 0) T = omp_get_max_threads();
  if (mode == "serial") T = 1;
  omp_set_num_threads(T);
  n = std::max<long long>(T, n / T * T);
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
  std::vector<Job> jobs;
  {
    std::string kind;
    std::vector<Function> fs = bench_workload(wl, &kind);
    const unsigned long long seed = kind == "cartpole" ? 1 : kind == "quad" ? 2 : kind == "rocket" ? 3 : kind == "mc" ? 4 : 5;
    for (const Function& f : fs) jobs.push_back(make_job(f, n, mode, T, seed, kind));
  }

  std::vector<double> secs;
  for (int r = 0; r < reps + warm; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    for (auto& j : jobs) run(j);
    auto t1 = std::chrono::steady_clock::now();
    if (r >= warm) secs.push_back(std::chrono::duration<double>(t1 - t0).count());
  }
  double total = 0;
  for (double v : secs) total += v;
  std::sort(secs.begin(), secs.end());
  const double med = secs[secs.size() / 2], best = secs.front();
  double chk = 0;
  for (auto& j : jobs) for (auto& o : j.out) for (double v : o) chk += v == v ? v : 0;
  printf("{\"workload\": \"%s\", \"n\": %lld, \"mode\": \"%s\", \"threads\": %d, \"reps\": %d, \"secs_median\": %.6f, "
         "\"secs_best\": %.6f, \"secs_total\": %.6f, \"evals_per_s\": %.6g, \"checksum\": %.17g}\n",
         wl.c_str(), n, mode.c_str(), T, reps, med, best, total, static_cast<double>(n) / med, chk);
  return 0;
}
