// C ABI of libcasadi_cuda.so (include/casadi_cuda.h): handle management, plan selection,
// host- and device-pointer evaluation paths.  No CPU fallback anywhere: every evaluation entry
// point fails when the tape has no CUDA device.
#include "../../include/casadi_cuda.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ccu_isa.h"
#include "comm.hpp"
#include "interp.cuh"
#include "jit.hpp"
#include "tape_reroll.hpp"
#include "layout.cuh"
#include "reduce.cuh"
#include "tape_builder.hpp"
#include "tape_compile.hpp"

namespace ccu { void host_copy_slice(void* dst, const void* src, size_t bytes, bool streaming); }  // hostcopy.cpp

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<int> g_default_mode{-1};  // -1: from the environment variable CCU_MODE

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CCU_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                   \
  } while (0)

struct DevBuf {
  double* p = nullptr;
  size_t cap = 0;  // doubles
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 1024;
    cudaError_t e = cudaMalloc(&p, want * sizeof(double));
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, n * sizeof(double));
      want = n;
    }
    if (e != cudaSuccess) return fail("cudaMalloc of %zu bytes failed: %s", n * sizeof(double), cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// page-locked host staging buffer
struct PinBuf {
  double* p = nullptr;
  size_t cap = 0;  // doubles
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaHostAlloc(&p, n * sizeof(double), cudaHostAllocDefault);
    if (e != cudaSuccess) { p = nullptr; return fail("cudaHostAlloc of %zu bytes failed: %s", n * sizeof(double), cudaGetErrorString(e)); }
    cap = n;
    return 0;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Multi-threaded memcpy between the caller's pageable buffers and the pinned staging of the host path: one thread
// moves ~10 GB/s, the PCIe link 55 GB/s, so the copies of a chunk are cut into slices taken by a few persistent
// workers (CCU_HOST_THREADS, default min(16, cores - 2)) and the calling thread.
class CopyPool {
 public:
  // never destroyed: the detached workers wait on cv_ for the life of the process, and destroying a condition variable
  // with waiters blocks (glibc) -- a function-local static object here hung every host program at exit
  static CopyPool& get() { static CopyPool* pool = new CopyPool; return *pool; }
  void copy(void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return;
    if (workers_.empty() || bytes <= kSlice) { ccu::host_copy_slice(dst, src, bytes, streaming_ && bytes >= (256u << 10)); return; }
    auto job = std::make_shared<Job>();
    job->dst = static_cast<char*>(dst); job->src = static_cast<const char*>(src); job->bytes = bytes;
    job->slices = (bytes + kSlice - 1) / kSlice;
    std::lock_guard<std::mutex> one(run_mu_);  // one copy at a time
    {
      std::lock_guard<std::mutex> lk(mu_);
      cur_ = job;
      ++epoch_;
    }
    cv_.notify_all();
    work(*job);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return job->done.load() == job->slices; });
  }
  int threads() const { return static_cast<int>(workers_.size()) + 1; }

 private:
  static constexpr size_t kSlice = 2u << 20;
  // the state of one copy; a worker that is late leaving a finished job only ever touches that job's counters
  struct Job {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0, slices = 0;
    std::atomic<size_t> next{0}, done{0};
  };
  CopyPool() {
    // all but two cores of this process' share of the machine (torchrun's LOCAL_WORLD_SIZE ranks divide it), at most 16:
    // 8 threads moved 37 GB/s on the 16-core B200 host, below the 50 GB/s of the PCIe link they feed
    int n = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) n /= std::max(1, atoi(lw));
    n = std::max(1, std::min(16, n - 2));
    if (const char* e = getenv("CCU_HOST_THREADS")) n = std::max(1, atoi(e));
    if (const char* e = getenv("CCU_HOST_STREAMING")) streaming_ = atoi(e) != 0;  // non-temporal stores (hostcopy.cpp), on by default
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    for (auto& t : workers_) t.detach();  // workers live as long as the process
  }
  void work(Job& j) {
    for (;;) {
      const size_t k = j.next.fetch_add(1);
      if (k >= j.slices) return;
      const size_t off = k * kSlice, len = std::min(kSlice, j.bytes - off);
      ccu::host_copy_slice(j.dst + off, j.src + off, len, streaming_);
      if (j.done.fetch_add(1) + 1 == j.slices) { std::lock_guard<std::mutex> lk(mu_); cv_done_.notify_all(); }
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        job = cur_;
      }
      if (job) work(*job);
    }
  }
  std::vector<std::thread> workers_;
  bool streaming_ = true;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, cv_done_;
  std::shared_ptr<Job> cur_;
  unsigned long long epoch_ = 0;
};

// true when the driver can DMA from / to `p` directly (page-locked or managed memory)
bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Page-locking the caller's buffers IN PLACE (cudaHostRegister), so that the DMA engines read and write them directly and
// the pinned staging (one extra read + write of every byte through the host memory system) drops out of the host path.
// Strictly opt-in -- ccu_host_register() or CCU_HOST_REGISTER=1 -- because a registration outlives a free() of the
// buffer: a caller that pins must keep the buffer alive until ccu_host_unregister() / the destruction of the tape that
// registered it (the reference's buffer API, Function::operator()(arg,res,iw,w,mem), is called with long-lived
// buffers; its DM API allocates fresh results per call and should not pin).
class HostRegistry {
 public:
  static HostRegistry& get() { static HostRegistry* r = new HostRegistry; return *r; }
  // 0: [p, p+bytes) is page-locked on return (by us, now or earlier, or by the caller); 1: it is not (stage it)
  int pin(const void* p, size_t bytes, const void* owner) {
    if (!p || bytes == 0) return 1;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    std::lock_guard<std::mutex> lk(mu_);
    auto it = map_.upper_bound(a);
    if (it != map_.begin()) {
      --it;
      if (it->first + it->second.bytes >= a + bytes) return 0;  // covered by an earlier registration
      if (it->first == a) {  // the same buffer grew: register it again in full
        cudaHostUnregister(reinterpret_cast<void*>(it->first));
        cudaGetLastError();
        map_.erase(it);
      }
    }
    if (host_ptr_is_pinned(p) && host_ptr_is_pinned(static_cast<const char*>(p) + bytes - 1)) return 0;  // the caller's own
    // a caller that presents new buffers at every call (the DM API) must not accumulate registrations: past a few dozen
    // per owner the automatic mode goes back to staging
    if (owner != nullptr) {
      size_t held = 0;
      for (const auto& e : map_) held += e.second.owner == owner;
      if (held >= kMaxPerOwner) return 1;
    }
    const cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return 1; }  // (overlaps a foreign registration, locked-memory limit, ...)
    map_[a] = Entry{bytes, owner};
    return 0;
  }
  int unpin(const void* p) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = map_.find(reinterpret_cast<uintptr_t>(p));
    if (it == map_.end()) return 1;
    const cudaError_t e = cudaHostUnregister(const_cast<void*>(p));
    cudaGetLastError();
    map_.erase(it);
    return e == cudaSuccess ? 0 : 1;
  }
  void release_owner(const void* owner) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = map_.begin(); it != map_.end();) {
      if (it->second.owner == owner) {
        cudaHostUnregister(reinterpret_cast<void*>(it->first));
        cudaGetLastError();
        it = map_.erase(it);
      } else {
        ++it;
      }
    }
  }
  size_t count() {
    std::lock_guard<std::mutex> lk(mu_);
    return map_.size();
  }

 private:
  static constexpr size_t kMaxPerOwner = 64;
  struct Entry { size_t bytes; const void* owner; };
  std::mutex mu_;
  std::map<uintptr_t, Entry> map_;
};

// CCU_HOST_REGISTER=1: page-lock every mapped caller buffer of at least 1 MiB the first time an evaluation sees it
bool host_register_auto() {
  const char* e = getenv("CCU_HOST_REGISTER");
  return e && atoi(e) != 0;
}

// staging of the host-pointer path (eval_host_impl): [slot] buffering
struct HostPipe {
  // Slots of the pipeline: chunk c is queued on the device, then the results of chunk c-(kSlots-1) are copied out of
  // the pinned staging.  A third slot (chunk c+1 queued before the copy-out of chunk c-1) was measured and does not pay:
  // with pageable caller buffers the host memory system is the limit -- the D2H DMA writes, the reads and the
  // non-temporal writes of the staging copy share it -- and the D2H phase simply stretches (quadrotor Jacobian, 2e6
  // instances: 50.1 ms with two slots, 52.0 ms with three; profiles/r2_e2e_staging.txt).
  static constexpr int kSlots = 2;
  bool ready = false;
  cudaStream_t s[3] = {nullptr, nullptr, nullptr};  // H2D, compute, D2H
  cudaEvent_t ev[kSlots][3] = {};                   // per slot: H2D done, compute done, D2H done
  std::vector<DevBuf> in_aos[kSlots], in_soa[kSlots], out_aos[kSlots], out_soa[kSlots];
  std::vector<DevBuf> bcast;  // reduce_in operands (one instance)
  DevBuf red;
  // pageable caller buffers go through pinned staging (the DMA engines cannot read pageable memory; the driver's own
  // staging of cudaMemcpyAsync is synchronous and serialises the pipeline)
  std::vector<PinBuf> in_pin[kSlots], out_pin[kSlots];
  PinBuf red_pin;
  cudaEvent_t tev[kSlots][6] = {};  // per slot: start/stop of the H2D, compute and D2H phase (timing enabled)
  bool tev_used[kSlots] = {};
  void release() {
    for (int b = 0; b < kSlots; ++b) {
      for (auto* v : {&in_aos[b], &in_soa[b], &out_aos[b], &out_soa[b]}) for (auto& x : *v) x.release();
      for (auto* v : {&in_pin[b], &out_pin[b]}) for (auto& x : *v) x.release();
      for (int k = 0; k < 3; ++k) if (ev[b][k]) cudaEventDestroy(ev[b][k]);
      for (int k = 0; k < 6; ++k) if (tev[b][k]) cudaEventDestroy(tev[b][k]);
    }
    red_pin.release();
    for (auto& x : bcast) x.release();
    red.release();
    for (int k = 0; k < 3; ++k) if (s[k]) cudaStreamDestroy(s[k]);
    ready = false;
  }
};

}  // namespace

struct ccu_tape {
  int device = -1;
  // source tape (kept for re-planning)
  std::vector<int> op, i0, i1, i2;
  std::vector<double> d;
  long long sz_w = 0;
  std::vector<long long> nnz_in, nnz_out;
  long long max_live = 0, flops = 0, cse_removed = 0;
  // compiled program + plan
  ccu::Program prog;
  ccu::LaunchPlan plan;
  bool use_acc = true;
  ccu::XInstr* d_prog = nullptr;
  long long n_records = 0;
  DevBuf scratch;
  // staging for the host-pointer path
  std::vector<DevBuf> d_in, d_out, d_part;
  DevBuf d_tmp;
  HostPipe pipe;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // last host-pointer evaluation: device time of the three pipeline phases (summed over the chunks; they overlap),
  // time the calling thread spent copying between pageable memory and the pinned staging, wall time, staged bytes
  double st_h2d_ms = 0, st_kernel_ms = 0, st_d2h_ms = 0, st_stage_ms = 0, st_wall_ms = 0, st_staged_bytes = 0;
  // specialised kernels (jit.hpp); mode: CCU_MODE_INTERP or CCU_MODE_JIT
  int mode = CCU_MODE_INTERP;
  ccu::JitOptions jit_opt;
  ccu::JitProgram jit;
  bool jit_built = false;
  std::string jit_error;
  mutable std::vector<std::string> jit_src;  // ccu_tape_get_jit_source: sources of the option set jit_src_key
  mutable std::string jit_src_key;
  int sms = 0;

  ccu::TapeSource source() const {
    ccu::TapeSource s;
    s.n_instr = static_cast<long long>(op.size());
    s.op = op.data(); s.i0 = i0.data(); s.i1 = i1.data(); s.i2 = i2.data(); s.d = d.data();
    s.sz_w = sz_w; s.nnz_in = nnz_in; s.nnz_out = nnz_out;
    return s;
  }
};

namespace {

// compile the program for (threads, ipt, S) and upload it
int build_plan(ccu_tape* t, int threads, int ipt, int slots_shared) {
  ccu::CompileOptions opt;
  opt.use_acc = t->use_acc;
  // automatic choice (B200 sweeps, profiles/r1_sweep_interp_v2.jsonl): keep the whole work vector in shared
  // memory when it is small; otherwise a SMALL shared window (occupancy beats fewer spills: 16 slots x 2 lanes
  // outran 48 x 1 by 1.9x on the quadrotor tape) and SPILL/FILL to the global scratch
  // The instruction order is the min-cut bisection order (tape_schedule.hpp) unless CCU_SCHED=0: it keeps far fewer
  // values alive at once than the reference's depth-first order, so more tapes fit the shared window entirely.
  opt.schedule = 1;
  if (const char* p = getenv("CCU_SCHED")) opt.schedule = atoi(p) ? 1 : 0;
  std::string err;
  ccu::Program prog;
  bool fits;
  if (slots_shared <= 0) {
    opt.slots_shared = 32;
    if (!ccu::compile_tape(t->source(), opt, &prog, &err)) return fail("tape compile failed: %s", err.c_str());
    fits = prog.spill_loads == 0 && prog.spill_stores == 0;
    if (!fits) {
      opt.slots_shared = 16;
      if (!ccu::compile_tape(t->source(), opt, &prog, &err)) return fail("tape compile failed: %s", err.c_str());
    }
  } else {
    opt.slots_shared = std::max(slots_shared, 4);
    if (!ccu::compile_tape(t->source(), opt, &prog, &err)) return fail("tape compile failed: %s", err.c_str());
    fits = prog.spill_loads == 0 && prog.spill_stores == 0;
  }
  const bool auto_threads = threads <= 0, auto_ipt = ipt <= 0;
  if (threads <= 0) threads = fits ? 256 : 128;
  if (ipt <= 0) ipt = 2;  // two lanes per thread amortise the dispatch and give the FP64 pipe independent work
  // a large window that holds the whole work vector: shrink the CTA until it fits the 227 KB
  auto window = [&] { return static_cast<size_t>(prog.slots_shared) * threads * ipt * 8; };
  if (auto_threads) while (window() > 200 * 1024 && threads > 64) threads /= 2;
  if (auto_ipt) while (window() > 200 * 1024 && ipt > 1) ipt /= 2;
  if (threads % 32 != 0 || threads > 1024) return fail("plan: threads must be a multiple of 32, <= 1024");
  if (ipt != 1 && ipt != 2 && ipt != 4) return fail("plan: ipt must be 1, 2 or 4");
  if (ipt == 4 && threads > 512) return fail("plan: at most 512 threads with 4 instances per thread");
  ccu::LaunchPlan plan;
  plan.threads = threads; plan.ipt = ipt;
  plan.slots_shared = prog.slots_shared; plan.slots_global = prog.slots_global;
  plan.smem_bytes = ccu::plan_smem_bytes(plan);
  if (plan.smem_bytes > 227 * 1024) return fail("plan needs %zu bytes of shared memory per CTA (> 227 KB)", plan.smem_bytes);
  if (t->device >= 0) {
    CCU_CUDA(cudaSetDevice(t->device));
    cudaError_t e = ccu::plan_occupancy(&plan, t->device);
    if (e != cudaSuccess) return fail("plan_occupancy: %s", cudaGetErrorString(e));
    if (t->d_prog) { cudaFree(t->d_prog); t->d_prog = nullptr; }
    std::vector<ccu::XInstr> rec = ccu::predecode(prog.words, threads * ipt);
    CCU_CUDA(cudaMalloc(&t->d_prog, rec.size() * sizeof(ccu::XInstr)));
    CCU_CUDA(cudaMemcpy(t->d_prog, rec.data(), rec.size() * sizeof(ccu::XInstr), cudaMemcpyHostToDevice));
    t->n_records = static_cast<long long>(rec.size());
  }
  t->prog = std::move(prog);
  t->plan = plan;
  return 0;
}

int ensure_scratch(ccu_tape* t) {
  if (t->plan.slots_global == 0) return 0;
  size_t n = static_cast<size_t>(t->plan.slots_global) * t->plan.grid * t->plan.threads * t->plan.ipt;
  return t->scratch.ensure(n);
}

int launch(ccu_tape* t, const ccu::IoDesc& io, long long N, cudaStream_t stream) {
  if (t->mode == CCU_MODE_JIT) {
    const long long tile = ccu::jit_tile_for(t->jit, N, t->sms);
    if (t->jit.chain) {
      if (t->jit.scratch_slots > 0 &&
          t->scratch.ensure(static_cast<size_t>(t->jit.scratch_slots) * t->jit.chain_grid * t->jit.threads)) return 1;
    } else if (t->jit.scratch_slots > 0 &&
               t->scratch.ensure(static_cast<size_t>(t->jit.scratch_slots) * tile * std::max(1, t->jit.streams))) return 1;
    if (t->ev0) cudaEventRecord(t->ev0, stream);
    long long nl = 0;
    cudaError_t e = ccu::jit_launch(t->jit, io, N, t->scratch.p, tile, stream, &nl);
    g_launches += nl;
    if (e != cudaSuccess) return fail("specialised kernel launch failed: %s", cudaGetErrorString(e));
    if (t->ev1) cudaEventRecord(t->ev1, stream);
    t->timed = true;
    return 0;
  }
  if (ensure_scratch(t)) return 1;
  if (t->ev0) cudaEventRecord(t->ev0, stream);
  cudaError_t e = ccu::launch_interp(t->plan, t->d_prog, t->n_records, io, N, t->scratch.p, stream);
  if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
  if (t->ev1) cudaEventRecord(t->ev1, stream);
  t->timed = true;
  if (N > 0) g_launches++;
  return 0;
}

void jit_options_from_env(ccu::JitOptions* o) {
  if (const char* p = getenv("CCU_JIT_SEG")) o->seg_instr = atoi(p);
  if (const char* p = getenv("CCU_JIT_SCHED")) o->schedule = atoi(p);
  if (const char* p = getenv("CCU_JIT_SEGWEIGHT")) o->seg_weight = atoll(p);
  if (const char* p = getenv("CCU_JIT_THREADS")) o->threads = atoi(p);
  if (const char* p = getenv("CCU_JIT_MINBLOCKS")) o->min_blocks = atoi(p);
  if (const char* p = getenv("CCU_JIT_BATCH")) o->load_batch = atoi(p);
  if (const char* p = getenv("CCU_JIT_TILE")) o->tile = atoll(p);
  if (const char* p = getenv("CCU_JIT_STREAMS")) o->streams = atoi(p);
  if (const char* p = getenv("CCU_JIT_PREFETCH")) o->prefetch = atoi(p);
  if (const char* p = getenv("CCU_JIT_SBLOCK")) o->scratch_block = atoi(p);
  if (const char* p = getenv("CCU_JIT_RING")) o->ring = atoi(p);
  if (const char* p = getenv("CCU_JIT_CHAIN")) o->chain = atoi(p);
  if (const char* p = getenv("CCU_JIT_INTERLEAVE")) o->interleave = atoi(p);
  if (const char* p = getenv("CCU_JIT_REMAT")) o->remat = atoi(p);
  if (const char* p = getenv("CCU_JIT_ROLL")) o->roll = atoi(p);
  if (const char* p = getenv("CCU_JIT_ROLL_REGS")) o->roll_registers = atoi(p);
  if (const char* p = getenv("CCU_JIT_SINCOS")) o->sincos = atoi(p);
  if (const char* p = getenv("CCU_JIT_FASTOPS")) o->fastops = atoi(p);
  if (const char* p = getenv("CCU_JIT_RING_INPUTS")) o->ring_inputs = atoi(p);
  if (const char* p = getenv("CCU_JIT_ZIGZAG")) o->zigzag = atoi(p);
  if (const char* p = getenv("CCU_JIT_IOBASE")) o->iobase = atoi(p);
  if (const char* p = getenv("CCU_JIT_STAGE")) o->stage = atoi(p);
  if (const char* p = getenv("CCU_JIT_SPILL")) o->spill = atoi(p);
  if (const char* p = getenv("CCU_JIT_REGVALS")) o->reg_values = atoi(p);
}

// (re)build the specialised kernels with the tape's current jit options
int build_jit(ccu_tape* t) {
  if (t->device < 0) return fail("tape was compiled without a CUDA device");
  const ccu::TapeSource tsrc = t->source();
  const ccu::JitOptions eff = ccu::jit_resolve(t->jit_opt, t->flops, &tsrc);
  if (eff.threads % 32 != 0 || eff.threads < 32 || eff.threads > 1024)
    return fail("jit: threads must be a multiple of 32 in [32, 1024]");
  CCU_CUDA(cudaSetDevice(t->device));
  ccu::JitProgram prog;
  std::string err;
  if (!ccu::jit_build(t->source(), eff, t->device, &prog, &err)) {
    t->jit_error = err;
    return fail("tape specialisation failed: %s", err.c_str());
  }
  if (prog.chain) {
    // persistent grid = what is resident at once
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(prog.chain), prog.threads,
                                                                 static_cast<size_t>(prog.chain_smem));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device);
    if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
    prog.chain_grid = per_sm * std::max(sms, 1);
  }
  if (t->jit_built) ccu::jit_destroy(&t->jit);
  t->jit = std::move(prog);
  t->jit_built = true;
  t->jit_error.clear();
  return 0;
}

int check_eval_args(const ccu_tape* t, long long N) {
  if (!t) return fail("null tape");
  if (t->device < 0) return fail("tape was compiled without a CUDA device: evaluation impossible (no CPU fallback)");
  if (N < 0) return fail("negative batch size");
  if (t->nnz_in.size() > ccu::kMaxIO || t->nnz_out.size() > ccu::kMaxIO)
    return fail("more than %d inputs or outputs are not supported", ccu::kMaxIO);
  return 0;
}

void fill_io(const ccu_tape* t, long long N, const double* const* d_arg, double* const* d_res, int layout,
             const int* reduce_in, ccu::IoDesc* io) {
  std::memset(io, 0, sizeof(*io));
  for (size_t j = 0; j < t->nnz_in.size(); ++j) {
    io->in[j] = t->nnz_in[j] > 0 ? d_arg[j] : nullptr;
    if (reduce_in && reduce_in[j]) { io->in_si[j] = 0; io->in_sk[j] = 1; }
    else if (layout == CCU_LAYOUT_SOA) { io->in_si[j] = 1; io->in_sk[j] = N; }
    else { io->in_si[j] = t->nnz_in[j]; io->in_sk[j] = 1; }
  }
  for (size_t j = 0; j < t->nnz_out.size(); ++j) {
    io->out[j] = t->nnz_out[j] > 0 ? d_res[j] : nullptr;
    if (layout == CCU_LAYOUT_SOA) { io->out_si[j] = 1; io->out_sk[j] = N; }
    else { io->out_si[j] = t->nnz_out[j]; io->out_sk[j] = 1; }
  }
}

}  // namespace

extern "C" {

int ccu_abi_version(void) { return CCU_ABI_VERSION; }
const char* ccu_last_error(void) { return g_err.c_str(); }

int ccu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

ccu_tape* ccu_tape_create(ccu_int n_instr, const int* op, const int* i0, const int* i1, const int* i2,
                          const double* d, ccu_int sz_w, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out,
                          const ccu_int* nnz_out, int device) {
  if (n_instr < 0 || n_in < 0 || n_out < 0 || (n_instr > 0 && (!op || !i0 || !i1 || !i2 || !d)) ||
      (n_in > 0 && !nnz_in) || (n_out > 0 && !nnz_out)) {
    fail("ccu_tape_create: invalid arguments");
    return nullptr;
  }
  ccu_tape* t = new ccu_tape();
  t->op.assign(op, op + n_instr); t->i0.assign(i0, i0 + n_instr); t->i1.assign(i1, i1 + n_instr);
  t->i2.assign(i2, i2 + n_instr); t->d.assign(d, d + n_instr);
  t->sz_w = sz_w;
  t->nnz_in.assign(nnz_in, nnz_in + n_in);
  t->nnz_out.assign(nnz_out, nnz_out + n_out);
  const char* noacc = getenv("CCU_NO_ACC");
  t->use_acc = !(noacc && noacc[0] == '1');
  std::string err;
  if (!ccu::analyse_tape(t->source(), &t->max_live, &t->flops, &err, &t->cse_removed)) {
    fail("invalid tape: %s", err.c_str());
    delete t;
    return nullptr;
  }
  if (device >= 0) {
    int n = ccu_device_count();
    if (device >= n) {
      fail("CUDA device %d not available (%d device(s) visible); there is no CPU fallback", device, n);
      delete t;
      return nullptr;
    }
    t->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&t->ev0) != cudaSuccess || cudaEventCreate(&t->ev1) != cudaSuccess) {
      fail("CUDA initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
      delete t;
      return nullptr;
    }
  }
  int threads = 0, ipt = 0, S = 0;
  if (const char* p = getenv("CCU_PLAN")) sscanf(p, "%d,%d,%d", &threads, &ipt, &S);
  if (build_plan(t, threads, ipt, S)) {
    ccu_tape_destroy(t);
    return nullptr;
  }
  t->d_in.resize(n_in); t->d_out.resize(n_out); t->d_part.resize(n_out);
  // execution mode: CCU_MODE = interp | jit | auto (default).  "auto" specialises the tape when NVRTC is
  // loadable and the compilation succeeds, and otherwise runs the interpreter kernel.
  jit_options_from_env(&t->jit_opt);
  if (t->device >= 0) {
    cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, t->device);
    const char* m = getenv("CCU_MODE");
    std::string want = m ? m : "auto";
    const int dm = g_default_mode.load();
    if (dm == CCU_MODE_INTERP) want = "interp";
    else if (dm == CCU_MODE_JIT) want = "jit";
    else if (dm == CCU_MODE_AUTO) want = "auto";
    if (want != "interp" && want != "jit" && want != "auto") {
      fail("CCU_MODE must be interp, jit or auto (got '%s')", want.c_str());
      ccu_tape_destroy(t);
      return nullptr;
    }
    if (want != "interp") {
      if (build_jit(t) == 0) {
        t->mode = CCU_MODE_JIT;
      } else if (want == "jit") {
        ccu_tape_destroy(t);
        return nullptr;
      }
    }
  }
  return t;
}

void ccu_tape_destroy(ccu_tape* t) {
  if (!t) return;
  HostRegistry::get().release_owner(t);
  if (t->device >= 0) {
    cudaSetDevice(t->device);
    if (t->d_prog) cudaFree(t->d_prog);
    if (t->jit_built) ccu::jit_destroy(&t->jit);
    t->scratch.release(); t->d_tmp.release();
    t->pipe.release();
    for (auto& b : t->d_in) b.release();
    for (auto& b : t->d_out) b.release();
    for (auto& b : t->d_part) b.release();
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->stream) cudaStreamDestroy(t->stream);
  }
  delete t;
}

int ccu_tape_get_info(const ccu_tape* t, ccu_tape_info* info) {
  if (!t || !info) return fail("null argument");
  std::memset(info, 0, sizeof(*info));
  info->n_instr = t->prog.n_instr;
  info->n_words = static_cast<ccu_int>(t->prog.words.size());
  info->flops = t->flops;
  for (auto v : t->nnz_in) info->bytes_in += 8 * v;
  for (auto v : t->nnz_out) info->bytes_out += 8 * v;
  info->sz_w = t->sz_w;
  info->slots_shared = t->plan.slots_shared;
  info->slots_global = t->plan.slots_global;
  info->threads = t->plan.threads;
  info->ipt = t->plan.ipt;
  info->smem_bytes = static_cast<ccu_int>(t->plan.smem_bytes);
  info->spill_loads = t->prog.spill_loads;
  info->spill_stores = t->prog.spill_stores;
  info->max_live = t->max_live;
  info->grid = t->plan.grid;
  info->ctas_per_sm = t->plan.ctas_per_sm;
  info->mode = t->mode;
  info->cse_removed = t->cse_removed;
  if (t->jit_built) {
    info->jit_segments = t->jit.segments;
    info->jit_chained = t->jit.chain ? 1 : 0;
    info->jit_remat_cloned = t->jit.remat_cloned;
    info->jit_loop_iters = t->jit.loop_iters;
    info->jit_loop_body = t->jit.loop_body;
    info->jit_loop_slots = t->jit.loop_slots;
    info->jit_scratch_slots = t->jit.scratch_slots;
    info->jit_tile = t->jit.tile;
    info->jit_compile_ms = static_cast<ccu_int>(t->jit.compile_ms);
    info->jit_cross_loads = t->jit.cross_loads;
    info->jit_cross_stores = t->jit.cross_stores;
    info->jit_max_regs = t->jit.max_regs;
    info->jit_cache_hits = t->jit.cache_hits;
    info->jit_threads = t->jit.threads;
    info->jit_schedule = t->jit_opt.schedule;
    info->jit_schedule_ms = static_cast<ccu_int>(t->jit.schedule_ms);
  }
  return 0;
}

int ccu_set_default_mode(int mode) {
  if (mode < -1 || mode > CCU_MODE_AUTO) return fail("unknown mode %d", mode);
  g_default_mode = mode;
  return 0;
}

int ccu_tape_set_mode(ccu_tape* t, int mode) {
  if (!t) return fail("null tape");
  if (mode == CCU_MODE_INTERP) { t->mode = CCU_MODE_INTERP; return 0; }
  if (mode != CCU_MODE_JIT) return fail("unknown mode %d", mode);
  if (!t->jit_built && build_jit(t)) return 1;
  t->mode = CCU_MODE_JIT;
  return 0;
}

int ccu_tape_set_jit_plan(ccu_tape* t, int seg_instr, int threads, int min_blocks, ccu_int tile) {
  if (!t) return fail("null tape");
  ccu::JitOptions o = t->jit_opt;
  if (seg_instr > 0) o.seg_instr = seg_instr;
  if (threads > 0) o.threads = threads;
  if (min_blocks >= 0) o.min_blocks = min_blocks;
  if (tile >= 0) o.tile = tile;
  std::swap(o, t->jit_opt);
  if (build_jit(t)) { std::swap(o, t->jit_opt); return 1; }
  t->mode = CCU_MODE_JIT;
  return 0;
}

int ccu_tape_jit_plan_stats(const ccu_tape* t, int seg_instr, int schedule, ccu_int stats[8]) {
  if (!t || !stats) return fail("null argument");
  ccu::JitOptions o = t->jit_opt;
  if (seg_instr > 0) o.seg_instr = seg_instr;
  if (schedule >= 0) o.schedule = schedule;
  const ccu::TapeSource tsrc = t->source();
  o = ccu::jit_resolve(o, t->flops, &tsrc);
  if (t->jit_opt.remat < 0) o.remat = 0;  // the order and the cuts alone; ccu_tape_jit_remat_stats reports the plan with recomputation
  ccu::JitPlanStats ps;
  std::string err;
  if (!ccu::jit_plan_stats(t->source(), o, &ps, &err)) return fail("%s", err.c_str());
  stats[0] = ps.segments; stats[1] = ps.scratch_slots; stats[2] = ps.cross_loads; stats[3] = ps.cross_stores;
  stats[4] = ps.max_segment; stats[5] = static_cast<ccu_int>(ps.schedule_ms); stats[6] = ps.max_live; stats[7] = static_cast<ccu_int>(ps.mean_live);
  return 0;
}

int ccu_tape_jit_remat_stats(const ccu_tape* t, int seg_instr, int remat, ccu_int stats[6]) {
  if (!t || !stats) return fail("null argument");
  ccu::JitOptions o = t->jit_opt;
  if (seg_instr > 0) o.seg_instr = seg_instr;
  if (remat >= 0) o.remat = remat;
  const ccu::TapeSource tsrc = t->source();
  o = ccu::jit_resolve(o, t->flops, &tsrc);
  ccu::JitPlanStats ps;
  std::string err;
  if (!ccu::jit_plan_stats(t->source(), o, &ps, &err)) return fail("%s", err.c_str());
  stats[0] = ps.remat_cloned; stats[1] = ps.remat_dropped; stats[2] = ps.cross_loads; stats[3] = ps.cross_stores;
  stats[4] = ps.segments; stats[5] = ps.scratch_slots;
  return 0;
}

int ccu_tape_loop_stats(const ccu_tape* t, ccu_int stats[8]) {
  if (!t || !stats) return fail("null argument");
  for (int k = 0; k < 8; ++k) stats[k] = 0;
  std::vector<ccu::Node> nodes;
  long long flops = 0;
  std::string err;
  if (!ccu::build_graph(t->source(), &nodes, &flops, &err)) return fail("%s", err.c_str());
  ccu::Roll roll;
  ccu::LoopTemplate tpl;
  if (!ccu::find_loop(nodes, &roll) || !ccu::build_loop_template(nodes, roll, &tpl, &roll.why)) {
    g_err = "no loop: " + roll.why;
    return 0;
  }
  long long before = 0;
  for (size_t v = 0; v < nodes.size(); ++v) before += nodes[v].kind == ccu::K_ARITH && roll.where[v] == -1;
  stats[0] = 1; stats[1] = tpl.K; stats[2] = tpl.B; stats[3] = static_cast<ccu_int>(tpl.carried.size());
  stats[4] = static_cast<ccu_int>(tpl.ctab.size()); stats[5] = static_cast<ccu_int>(tpl.affine.size());
  stats[6] = static_cast<ccu_int>(tpl.exit_pos.size()); stats[7] = before;
  return 0;
}

int ccu_tape_set_jit_remat(ccu_tape* t, int remat) {
  if (!t) return fail("null tape");
  const int old = t->jit_opt.remat;
  t->jit_opt.remat = remat;
  if (build_jit(t)) { t->jit_opt.remat = old; return 1; }
  t->mode = CCU_MODE_JIT;
  return 0;
}

ccu_int ccu_tape_jit_link_check(const ccu_tape* t) {
  if (!t) { fail("null tape"); return -1; }
  const ccu::TapeSource tsrc = t->source();
  const ccu::JitOptions eff = ccu::jit_resolve(t->jit_opt, t->flops, &tsrc);
  std::vector<std::string> src;
  std::string err, image;
  if (!ccu::jit_generate(t->source(), eff, &src, nullptr, &err)) { fail("%s", err.c_str()); return -1; }
  if (!ccu::jit_link_chain(src, eff, "sm_100a", &image, nullptr, &err)) { fail("%s", err.c_str()); return -1; }
  return static_cast<ccu_int>(image.size());
}

ccu_int ccu_tape_jit_compile_check(const ccu_tape* t, const char* dump_dir) {
  if (!t) { fail("null tape"); return -1; }
  const ccu::TapeSource tsrc = t->source();
  const ccu::JitOptions eff = ccu::jit_resolve(t->jit_opt, t->flops, &tsrc);
  std::string err;
  const long long n = ccu::jit_compile_check(t->source(), eff, "sm_100a", dump_dir ? dump_dir : "", &err);
  if (n < 0) fail("%s", err.c_str());
  return n;
}

const char* ccu_tape_jit_chain_error(const ccu_tape* t) {
  return (t && t->jit_built) ? t->jit.chain_error.c_str() : "";
}

int ccu_tape_set_jit_schedule(ccu_tape* t, int schedule) {
  if (!t) return fail("null tape");
  if (schedule < 0 || schedule > 1) return fail("unknown schedule %d", schedule);
  t->jit_opt.schedule = schedule;
  return 0;
}

ccu_int ccu_tape_get_jit_source(const ccu_tape* t, ccu_int segment, char* buf, ccu_int cap) {
  if (!t) { fail("null tape"); return -1; }
  // generating re-plans the whole tape: keep the sources of the last option set
  char key[160];
  const ccu::TapeSource tsrc = t->source();
  const ccu::JitOptions eff = ccu::jit_resolve(t->jit_opt, t->flops, &tsrc);
  snprintf(key, sizeof key, "%d,%d,%d,%d,%lld,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d", eff.fastops, eff.sincos, eff.interleave, eff.ring_inputs, eff.seg_weight, eff.seg_instr, eff.schedule, eff.threads, eff.min_blocks,
           eff.load_batch, eff.stage, eff.spill, eff.reg_values, eff.prefetch, eff.scratch_block, eff.ring);
  if (t->jit_src_key != key) {
    std::string err;
    t->jit_src.clear();
    t->jit_src_key.clear();
    if (!ccu::jit_generate(t->source(), eff, &t->jit_src, nullptr, &err)) { fail("%s", err.c_str()); return -1; }
    t->jit_src_key = key;
  }
  const std::vector<std::string>& src = t->jit_src;
  if (segment < 0) return static_cast<ccu_int>(src.size());
  if (segment >= static_cast<ccu_int>(src.size())) { fail("segment out of range"); return -1; }
  const std::string& s = src[segment];
  if (buf && cap > 0) {
    size_t n = std::min<size_t>(s.size(), static_cast<size_t>(cap) - 1);
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return static_cast<ccu_int>(s.size());
}

int ccu_tape_set_plan(ccu_tape* t, int threads, int ipt, int slots_shared) {
  if (!t) return fail("null tape");
  return build_plan(t, threads, ipt, slots_shared);
}

ccu_int ccu_tape_get_program(const ccu_tape* t, unsigned long long* words, ccu_int cap) {
  if (!t) { fail("null tape"); return -1; }
  ccu_int n = static_cast<ccu_int>(t->prog.words.size());
  if (words) std::memcpy(words, t->prog.words.data(), 8 * static_cast<size_t>(n < cap ? n : cap));
  return n;
}

int ccu_map_eval_device(ccu_tape* t, ccu_int N, const double* const* d_arg, double* const* d_res, int layout,
                        void* stream) {
  if (check_eval_args(t, N)) return 1;
  if (layout != CCU_LAYOUT_AOS && layout != CCU_LAYOUT_SOA) return fail("unknown layout %d", layout);
  if (N == 0) return 0;
  CCU_CUDA(cudaSetDevice(t->device));
  ccu::IoDesc io;
  fill_io(t, N, d_arg, d_res, layout, nullptr, &io);
  return launch(t, io, N, static_cast<cudaStream_t>(stream));
}

// shared by ccu_map_eval_reduce_device (part_out == NULL: block sums + tree into d_res[j]) and
// ccu_map_eval_shard_device (block sums of the shard written to part_out[j], no tree)
static int eval_reduce_core(ccu_tape* t, ccu_int N, const double* const* d_arg, double* const* d_res,
                            const int* reduce_in, const int* reduce_out, int layout, cudaStream_t stream,
                            double* const* part_out) {
  if (check_eval_args(t, N)) return 1;
  if (layout != CCU_LAYOUT_AOS && layout != CCU_LAYOUT_SOA) return fail("unknown layout %d", layout);
  CCU_CUDA(cudaSetDevice(t->device));
  const size_t n_out = t->nnz_out.size();
  auto reduced = [&](size_t j) {
    return reduce_out && reduce_out[j] && t->nnz_out[j] > 0 && (part_out ? part_out[j] != nullptr : d_res[j] != nullptr);
  };
  // reduced outputs are first evaluated per instance into a temporary, then tree-summed
  std::vector<double*> res(n_out);
  for (size_t j = 0; j < n_out; ++j) {
    res[j] = d_res ? d_res[j] : nullptr;
    if (reduce_out && reduce_out[j]) res[j] = nullptr;
    if (reduced(j)) {
      if (t->d_out[j].ensure(static_cast<size_t>(std::max<ccu_int>(N, 1)) * t->nnz_out[j])) return 1;
      res[j] = t->d_out[j].p;
    }
  }
  ccu::IoDesc io;
  fill_io(t, N, d_arg, res.data(), layout, reduce_in, &io);
  if (N > 0 && launch(t, io, N, stream)) return 1;
  const long long nblocks = (N + ccu::kReduceBlock - 1) / ccu::kReduceBlock;
  for (size_t j = 0; j < n_out; ++j) {
    if (!reduced(j)) continue;
    const int nnz = static_cast<int>(t->nnz_out[j]);
    if (part_out) {
      CCU_CUDA(ccu::launch_block_sums(res[j], io.out_si[j], io.out_sk[j], N, nnz, part_out[j], stream));
      g_launches += (N > 0 ? 1 : 0);
    } else {
      if (t->d_part[j].ensure(static_cast<size_t>(nblocks > 0 ? nblocks : 1) * nnz)) return 1;
      CCU_CUDA(ccu::launch_block_sums(res[j], io.out_si[j], io.out_sk[j], N, nnz, t->d_part[j].p, stream));
      CCU_CUDA(ccu::launch_tree(t->d_part[j].p, nblocks, nnz, d_res[j], stream));
      g_launches += (N > 0 ? 2 : 1);
    }
  }
  return 0;
}

int ccu_map_eval_reduce_device(ccu_tape* t, ccu_int N, const double* const* d_arg, double* const* d_res,
                               const int* reduce_in, const int* reduce_out, int layout, void* stream) {
  return eval_reduce_core(t, N, d_arg, d_res, reduce_in, reduce_out, layout, static_cast<cudaStream_t>(stream), nullptr);
}

int ccu_map_eval_shard_device(ccu_tape* t, ccu_int N_global, ccu_int i0, ccu_int n, const double* const* d_arg,
                              double* const* d_res, const int* reduce_in, const int* reduce_out, double* const* d_part,
                              int layout, void* stream) {
  if (!t) return fail("null tape");
  if (i0 < 0 || n < 0 || i0 + n > N_global) return fail("shard [%lld, %lld) outside the batch of %lld", i0, i0 + n, N_global);
  if (n > 0 && i0 % ccu::kReduceBlock != 0) return fail("shard offset %lld is not a multiple of the reduction block %d", i0, ccu::kReduceBlock);
  const size_t n_out = t->nnz_out.size();
  std::vector<double*> part(n_out, nullptr);
  bool any = false;
  for (size_t j = 0; j < n_out; ++j) {
    if (reduce_out && reduce_out[j] && d_part && d_part[j]) {
      part[j] = d_part[j] + (i0 / ccu::kReduceBlock) * t->nnz_out[j];
      any = true;
    }
  }
  if (!any && !(reduce_in))
    return ccu_map_eval_device(t, n, d_arg, d_res, layout, stream);
  return eval_reduce_core(t, n, d_arg, d_res, reduce_in, reduce_out, layout, static_cast<cudaStream_t>(stream),
                          any ? part.data() : nullptr);
}

int ccu_reduce_tree_device(int device, double* d_part, ccu_int N_global, ccu_int nnz, double* d_out, void* stream) {
  if (!d_part || !d_out || N_global < 0 || nnz < 0) return fail("ccu_reduce_tree_device: invalid arguments");
  CCU_CUDA(cudaSetDevice(device));
  const long long nblocks = (N_global + ccu::kReduceBlock - 1) / ccu::kReduceBlock;
  CCU_CUDA(ccu::launch_tree(d_part, nblocks, static_cast<int>(nnz), d_out, static_cast<cudaStream_t>(stream)));
  g_launches++;
  return 0;
}

// Host-pointer evaluation (what CudaMap::eval calls): the batch is cut into chunks that flow through a
// pipeline on three streams with double-buffered device staging,
//     [pageable -> pinned]  H2D (AoS)  ->  [AoS->SoA, tape kernels on SoA, SoA->AoS | block sums]  ->  D2H (AoS)  [pinned -> pageable]
// so the PCIe transfers of chunk c+1 / c-1 overlap the kernels of chunk c, and the tape kernels always run
// on the coalesced SoA layout.  Caller buffers that are not page-locked (what a CasADi caller normally owns:
// std::vector / DM storage) are copied through pinned staging by the calling thread and the CopyPool workers while
// the device works on the neighbouring chunks; page-locked buffers are used directly.  Chunks are multiples of
// kReduceBlock, so reduce_out block sums land at their global positions and the summation tree is the same as for a
// single launch.
// (g_off, N_glob): position of these N instances inside the whole batch when the batch is sharded over several devices
// (ccu_multi); the block sums of reduced outputs are then written at their GLOBAL rows of d_part and the cross-device
// combine + level-1 tree is left to the caller (finish_reduce == false).
// in_groups / out_groups (may be NULL): input / output j of an instance is G_j pieces of nnz/G_j doubles and the caller's
// buffer is PIECE-major -- piece d of instance k at arg[j] + (d*N_glob + k) * (nnz/G_j) -- the layout in which the
// reference hands the nfwd (nadj) seed / sensitivity blocks to a derivative map (map.cpp:231-264, 285-318: the
// GetNonzeros column permutations around df.map(n)); the permutation is folded into the chunk copies here.  For grouped
// buffers arg[j] / res[j] are the UNSHIFTED bases even for a shard (g_off is added here).
static int eval_host_chunks(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res,
                            const int* reduce_in, const int* reduce_out, long long g_off, long long N_glob,
                            bool finish_reduce, const int* in_groups, const int* out_groups) {
  const size_t n_in = t->nnz_in.size(), n_out = t->nnz_out.size();
  HostPipe& hp = t->pipe;
  cudaStream_t s_h2d = hp.s[0], s_cmp = hp.s[1], s_d2h = hp.s[2];
  // chunk size: a multiple of kReduceBlock; ~8 chunks per call, between 64Ki and 512Ki instances
  long long C = (N + 7) / 8;
  if (const char* p = getenv("CCU_HOST_CHUNK")) C = atoll(p);
  else C = std::max<long long>(1 << 16, std::min<long long>(1 << 19, C));  // (quad_ms, 4e6 instances: 256-512 Ki chunks +4 % over 1 Mi)
  C = std::max<long long>(ccu::kReduceBlock, (C + ccu::kReduceBlock - 1) / ccu::kReduceBlock * ccu::kReduceBlock);
  const long long nchunks = N > 0 ? (N + C - 1) / C : 0;
  const long long nblocks = (N_glob + ccu::kReduceBlock - 1) / ccu::kReduceBlock;
  const long long blk_off = g_off / ccu::kReduceBlock;
  auto is_rin = [&](size_t j) { return reduce_in && reduce_in[j]; };
  auto is_rout = [&](size_t j) { return reduce_out && reduce_out[j]; };
  auto gin = [&](size_t j) { return (in_groups && in_groups[j] > 1) ? in_groups[j] : 1; };
  auto gout = [&](size_t j) { return (out_groups && out_groups[j] > 1) ? out_groups[j] : 1; };
  for (size_t j = 0; j < n_in; ++j)
    if (gin(j) > 1 && (t->nnz_in[j] % gin(j) != 0 || is_rin(j))) return fail("input %zu: invalid piece count %d", j, gin(j));
  for (size_t j = 0; j < n_out; ++j)
    if (gout(j) > 1 && (t->nnz_out[j] % gout(j) != 0 || is_rout(j))) return fail("output %zu: invalid piece count %d", j, gout(j));
  // caller address of piece d of the chunk starting at local instance i0 (ungrouped: d = 0, the shard-shifted pointer)
  auto in_src = [&](size_t j, int d, long long i0) {
    const long long m = t->nnz_in[j] / gin(j);
    return gin(j) > 1 ? arg[j] + (d * N_glob + g_off + i0) * m : arg[j] + i0 * m;
  };
  auto out_dst = [&](size_t j, int d, long long i0) {
    const long long m = t->nnz_out[j] / gout(j);
    return gout(j) > 1 ? res[j] + (d * N_glob + g_off + i0) * m : res[j] + i0 * m;
  };
  // which caller buffers need pinned staging
  std::vector<char> stage_in(n_in, 0), stage_out(n_out, 0);
  const char* force = getenv("CCU_HOST_STAGING");  // "0": never stage (driver-staged copies), "1": always
  for (size_t j = 0; j < n_in; ++j)
    if (arg[j] && t->nnz_in[j] > 0 && !is_rin(j)) stage_in[j] = force ? force[0] == '1' : !host_ptr_is_pinned(arg[j]);
  for (size_t j = 0; j < n_out; ++j)
    if (res[j] && t->nnz_out[j] > 0 && !is_rout(j)) stage_out[j] = force ? force[0] == '1' : !host_ptr_is_pinned(res[j]);
  auto stage_ms = [&](const std::chrono::steady_clock::time_point& t0) {
    t->st_stage_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  CopyPool& pool = CopyPool::get();
  // broadcast (reduce_in) inputs: one instance, copied once
  for (size_t j = 0; j < n_in; ++j) {
    if (!is_rin(j) || !arg[j] || t->nnz_in[j] == 0) continue;
    if (hp.bcast[j].ensure(static_cast<size_t>(t->nnz_in[j]))) return 1;
    CCU_CUDA(cudaMemcpyAsync(hp.bcast[j].p, arg[j], t->nnz_in[j] * 8, cudaMemcpyHostToDevice, s_cmp));
  }
  for (size_t j = 0; j < n_out; ++j) {
    if (!is_rout(j) || !res[j] || t->nnz_out[j] == 0) continue;
    if (t->d_part[j].ensure(static_cast<size_t>(std::max<long long>(nblocks, 1)) * t->nnz_out[j])) return 1;
    // a shard owns only its rows: the others must read as zero for the cross-device combine
    if (!finish_reduce)
      CCU_CUDA(cudaMemsetAsync(t->d_part[j].p, 0, static_cast<size_t>(std::max<long long>(nblocks, 1)) * t->nnz_out[j] * 8, s_cmp));
  }
  // accumulate the phase times of the chunk that used slot b last
  auto harvest = [&](int b) {
    if (!hp.tev_used[b]) return;
    float f = 0;
    if (cudaEventSynchronize(hp.tev[b][5]) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaEventElapsedTime(&f, hp.tev[b][0], hp.tev[b][1]) == cudaSuccess) t->st_h2d_ms += f;
    if (cudaEventElapsedTime(&f, hp.tev[b][2], hp.tev[b][3]) == cudaSuccess) t->st_kernel_ms += f;
    if (cudaEventElapsedTime(&f, hp.tev[b][4], hp.tev[b][5]) == cudaSuccess) t->st_d2h_ms += f;
    cudaGetLastError();
    hp.tev_used[b] = false;
  };
  // copy the results of chunk c out of the pinned staging of its slot (after its D2H has completed)
  auto copy_out = [&](long long c) -> int {
    const int b = static_cast<int>(c % HostPipe::kSlots);
    const long long i0 = c * C, n = std::min(C, N - i0);
    bool any = false;
    for (size_t j = 0; j < n_out; ++j) any = any || stage_out[j];
    if (!any) return 0;
    CCU_CUDA(cudaEventSynchronize(hp.ev[b][2]));
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t j = 0; j < n_out; ++j) {
      if (!stage_out[j]) continue;
      const long long m = t->nnz_out[j] / gout(j);
      for (int d = 0; d < gout(j); ++d) {
        const size_t bytes = static_cast<size_t>(n) * m * 8;
        pool.copy(out_dst(j, d, i0), hp.out_pin[b][j].p + static_cast<size_t>(d) * n * m, bytes);
        t->st_staged_bytes += static_cast<double>(bytes);
      }
    }
    stage_ms(t0);
    return 0;
  };
  for (long long c = 0; c < nchunks; ++c) {
    const int b = static_cast<int>(c % HostPipe::kSlots);
    const long long i0 = c * C, n = std::min(C, N - i0);
    std::vector<const double*> d_arg(n_in, nullptr);
    std::vector<double*> d_res(n_out, nullptr);
    harvest(b);
    // ---- stage 0 (host): pageable inputs of chunk c -> pinned staging of slot b (free once the H2D of chunk c-2 is done)
    {
      bool any = false;
      for (size_t j = 0; j < n_in; ++j) any = any || stage_in[j];
      if (any) {
        if (c >= HostPipe::kSlots) CCU_CUDA(cudaEventSynchronize(hp.ev[b][0]));
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t j = 0; j < n_in; ++j) {
          if (!stage_in[j]) continue;
          if (hp.in_pin[b][j].ensure(static_cast<size_t>(t->nnz_in[j]) * C)) return 1;
          const long long m = t->nnz_in[j] / gin(j);
          for (int d = 0; d < gin(j); ++d) {
            const size_t bytes = static_cast<size_t>(n) * m * 8;
            pool.copy(hp.in_pin[b][j].p + static_cast<size_t>(d) * n * m, in_src(j, d, i0), bytes);
            t->st_staged_bytes += static_cast<double>(bytes);
          }
        }
        stage_ms(t0);
      }
    }
    // ---- stage 1: H2D (the device staging of slot b is free once chunk c-2 has been computed)
    CCU_CUDA(cudaStreamWaitEvent(s_h2d, hp.ev[b][1], 0));
    CCU_CUDA(cudaEventRecord(hp.tev[b][0], s_h2d));
    for (size_t j = 0; j < n_in; ++j) {
      if (!arg[j] || t->nnz_in[j] == 0) continue;
      if (is_rin(j)) { d_arg[j] = hp.bcast[j].p; continue; }
      const size_t cnt = static_cast<size_t>(t->nnz_in[j]) * C;
      if (hp.in_aos[b][j].ensure(cnt) || hp.in_soa[b][j].ensure(cnt)) return 1;
      if (stage_in[j] || gin(j) == 1) {
        const double* src = stage_in[j] ? hp.in_pin[b][j].p : in_src(j, 0, i0);
        CCU_CUDA(cudaMemcpyAsync(hp.in_aos[b][j].p, src, static_cast<size_t>(n) * t->nnz_in[j] * 8, cudaMemcpyHostToDevice, s_h2d));
      } else {
        const long long m = t->nnz_in[j] / gin(j);
        for (int d = 0; d < gin(j); ++d)
          CCU_CUDA(cudaMemcpyAsync(hp.in_aos[b][j].p + static_cast<size_t>(d) * n * m, in_src(j, d, i0), static_cast<size_t>(n) * m * 8,
                                   cudaMemcpyHostToDevice, s_h2d));
      }
      d_arg[j] = hp.in_soa[b][j].p;
    }
    CCU_CUDA(cudaEventRecord(hp.tev[b][1], s_h2d));
    CCU_CUDA(cudaEventRecord(hp.ev[b][0], s_h2d));
    // ---- stage 2: compute (the output staging of slot b is free once chunk c-2 has been copied back)
    CCU_CUDA(cudaStreamWaitEvent(s_cmp, hp.ev[b][0], 0));
    CCU_CUDA(cudaStreamWaitEvent(s_cmp, hp.ev[b][2], 0));
    CCU_CUDA(cudaEventRecord(hp.tev[b][2], s_cmp));
    for (size_t j = 0; j < n_in; ++j) {
      if (!arg[j] || t->nnz_in[j] == 0 || is_rin(j)) continue;
      // (piece d of a grouped input is an AoS block of its own: rows [d*m, (d+1)*m) of the SoA operand)
      const long long m = t->nnz_in[j] / gin(j);
      for (int d = 0; d < gin(j); ++d) {
        CCU_CUDA(ccu::launch_aos_to_soa(hp.in_aos[b][j].p + static_cast<size_t>(d) * n * m, hp.in_soa[b][j].p + static_cast<size_t>(d) * m * n,
                                        n, static_cast<int>(m), n, s_cmp));
        g_launches++;
      }
    }
    for (size_t j = 0; j < n_out; ++j) {
      if (!res[j] || t->nnz_out[j] == 0) continue;
      const size_t cnt = static_cast<size_t>(t->nnz_out[j]) * C;
      if (hp.out_soa[b][j].ensure(cnt)) return 1;
      if (!is_rout(j) && hp.out_aos[b][j].ensure(cnt)) return 1;
      if (stage_out[j] && hp.out_pin[b][j].ensure(cnt)) return 1;
      d_res[j] = hp.out_soa[b][j].p;
    }
    ccu::IoDesc io;
    fill_io(t, n, d_arg.data(), d_res.data(), CCU_LAYOUT_SOA, reduce_in, &io);
    if (launch(t, io, n, s_cmp)) return 1;
    for (size_t j = 0; j < n_out; ++j) {
      if (!d_res[j]) continue;
      const int nnz = static_cast<int>(t->nnz_out[j]);
      if (is_rout(j)) {
        CCU_CUDA(ccu::launch_block_sums(d_res[j], 1, n, n, nnz, t->d_part[j].p + (blk_off + i0 / ccu::kReduceBlock) * nnz, s_cmp));
        g_launches++;
      } else {
        const long long m = nnz / gout(j);
        for (int d = 0; d < gout(j); ++d) {
          CCU_CUDA(ccu::launch_soa_to_aos(d_res[j] + static_cast<size_t>(d) * m * n, hp.out_aos[b][j].p + static_cast<size_t>(d) * n * m, n,
                                          static_cast<int>(m), n, s_cmp));
          g_launches++;
        }
      }
    }
    CCU_CUDA(cudaEventRecord(hp.tev[b][3], s_cmp));
    CCU_CUDA(cudaEventRecord(hp.ev[b][1], s_cmp));
    // ---- stage 3: D2H
    CCU_CUDA(cudaStreamWaitEvent(s_d2h, hp.ev[b][1], 0));
    CCU_CUDA(cudaEventRecord(hp.tev[b][4], s_d2h));
    for (size_t j = 0; j < n_out; ++j) {
      if (!d_res[j] || is_rout(j)) continue;
      if (stage_out[j] || gout(j) == 1) {
        double* dst = stage_out[j] ? hp.out_pin[b][j].p : out_dst(j, 0, i0);
        CCU_CUDA(cudaMemcpyAsync(dst, hp.out_aos[b][j].p, static_cast<size_t>(n) * t->nnz_out[j] * 8, cudaMemcpyDeviceToHost, s_d2h));
      } else {
        const long long m = t->nnz_out[j] / gout(j);
        for (int d = 0; d < gout(j); ++d)
          CCU_CUDA(cudaMemcpyAsync(out_dst(j, d, i0), hp.out_aos[b][j].p + static_cast<size_t>(d) * n * m, static_cast<size_t>(n) * m * 8,
                                   cudaMemcpyDeviceToHost, s_d2h));
      }
    }
    CCU_CUDA(cudaEventRecord(hp.tev[b][5], s_d2h));
    CCU_CUDA(cudaEventRecord(hp.ev[b][2], s_d2h));
    hp.tev_used[b] = true;
    // ---- stage 4 (host): results of chunk c-2, pinned staging -> caller memory, while the device works on chunks c-1
    // and c (the output staging of slot b is rewritten by chunk c+3, issued after the copy-out of chunk c)
    if (c >= HostPipe::kSlots - 1 && copy_out(c - (HostPipe::kSlots - 1))) return 1;
  }
  for (long long c = std::max<long long>(0, nchunks - (HostPipe::kSlots - 1)); c < nchunks; ++c)
    if (copy_out(c)) return 1;
  // reduced outputs: level-1 tree over the block sums, then a tiny D2H
  for (size_t j = 0; j < n_out && finish_reduce; ++j) {
    if (!is_rout(j) || !res[j] || t->nnz_out[j] == 0) continue;
    const int nnz = static_cast<int>(t->nnz_out[j]);
    if (hp.red.ensure(static_cast<size_t>(nnz)) || hp.red_pin.ensure(static_cast<size_t>(nnz))) return 1;
    CCU_CUDA(ccu::launch_tree(t->d_part[j].p, nblocks, nnz, hp.red.p, s_cmp));
    g_launches++;
    CCU_CUDA(cudaMemcpyAsync(hp.red_pin.p, hp.red.p, static_cast<size_t>(nnz) * 8, cudaMemcpyDeviceToHost, s_cmp));
    CCU_CUDA(cudaStreamSynchronize(s_cmp));  // hp.red is reused by the next reduced output
    std::memcpy(res[j], hp.red_pin.p, static_cast<size_t>(nnz) * 8);
  }
  for (int b = 0; b < HostPipe::kSlots; ++b) harvest(b);
  return 0;
}

// Opt-in page-locking of the whole-batch caller buffers (HostRegistry) before the chunked pipeline looks at them: a
// buffer that is page-locked afterwards is copied by the DMA engines directly (host_ptr_is_pinned in eval_host_chunks).
static void register_caller_buffers(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res,
                                    const int* reduce_in, const int* reduce_out) {
  if (!host_register_auto() || N <= 0) return;
  const size_t kMin = size_t(1) << 20;
  HostRegistry& reg = HostRegistry::get();
  for (size_t j = 0; j < t->nnz_in.size(); ++j) {
    const size_t bytes = static_cast<size_t>(N) * t->nnz_in[j] * 8;
    if (arg[j] && bytes >= kMin && !(reduce_in && reduce_in[j])) reg.pin(arg[j], bytes, t);
  }
  for (size_t j = 0; j < t->nnz_out.size(); ++j) {
    const size_t bytes = static_cast<size_t>(N) * t->nnz_out[j] * 8;
    if (res[j] && bytes >= kMin && !(reduce_out && reduce_out[j])) reg.pin(res[j], bytes, t);
  }
}

static int eval_host_impl(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res,
                          const int* reduce_in, const int* reduce_out, long long g_off = 0, long long N_glob = -1,
                          bool finish_reduce = true, const int* in_groups = nullptr, const int* out_groups = nullptr) {
  if (N_glob < 0) N_glob = N;
  if (check_eval_args(t, N)) return 1;
  if (!arg || !res) return fail("null argument / result array");
  CCU_CUDA(cudaSetDevice(t->device));
  const size_t n_in = t->nnz_in.size(), n_out = t->nnz_out.size();
  HostPipe& hp = t->pipe;
  if (!hp.ready) {
    for (int k = 0; k < 3; ++k) CCU_CUDA(cudaStreamCreateWithFlags(&hp.s[k], cudaStreamNonBlocking));
    for (int b = 0; b < HostPipe::kSlots; ++b) {
      for (int k = 0; k < 3; ++k) CCU_CUDA(cudaEventCreateWithFlags(&hp.ev[b][k], cudaEventDisableTiming));
      for (int k = 0; k < 6; ++k) CCU_CUDA(cudaEventCreate(&hp.tev[b][k]));
    }
    for (int b = 0; b < HostPipe::kSlots; ++b) {
      hp.in_aos[b].resize(n_in); hp.in_soa[b].resize(n_in); hp.out_aos[b].resize(n_out); hp.out_soa[b].resize(n_out);
      hp.in_pin[b].resize(n_in); hp.out_pin[b].resize(n_out);
    }
    hp.bcast.resize(n_in);
    hp.ready = true;
  }
  if (g_off == 0 && N_glob == N) register_caller_buffers(t, N, arg, res, reduce_in, reduce_out);  // (a shard: done by the caller)
  t->st_h2d_ms = t->st_kernel_ms = t->st_d2h_ms = t->st_stage_ms = t->st_wall_ms = t->st_staged_bytes = 0;
  for (int b = 0; b < HostPipe::kSlots; ++b) hp.tev_used[b] = false;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = eval_host_chunks(t, N, arg, res, reduce_in, reduce_out, g_off, N_glob, finish_reduce, in_groups, out_groups);
  const std::string first_error = rc ? g_err : std::string();
  // Whatever happened, nothing may still read the caller's inputs or write the caller's outputs after the return
  cudaError_t e0 = cudaStreamSynchronize(hp.s[0]), e1 = cudaStreamSynchronize(hp.s[1]), e2 = cudaStreamSynchronize(hp.s[2]);
  cudaError_t e = e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2);
  if (rc == 0 && e != cudaSuccess) rc = fail("evaluation failed on device: %s", cudaGetErrorString(e));
  else if (rc) { g_err = first_error; cudaGetLastError(); }
  t->st_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int ccu_map_eval_host(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res) {
  return eval_host_impl(t, N, arg, res, nullptr, nullptr);
}

int ccu_map_eval_reduce_host(ccu_tape* t, ccu_int N, const double* const* arg, double* const* res,
                             const int* reduce_in, const int* reduce_out) {
  return eval_host_impl(t, N, arg, res, reduce_in, reduce_out);
}

// ------------------------------------------------------------------------------- communicators (comm.cu)
struct ccu_comm {
  ccu::Comm* c = nullptr;
};

int ccu_comm_available(void) {
  std::string why;
  if (ccu::comm_available(&why)) return 1;
  fail("%s", why.c_str());
  return 0;
}

int ccu_comm_unique_id(unsigned char id[128]) {
  std::string err;
  if (!id) return fail("null id");
  if (!ccu::comm_unique_id(id, &err)) return fail("%s", err.c_str());
  return 0;
}

ccu_comm* ccu_comm_create_all(int n_devices, const int* devices) {
  if (n_devices < 1) { fail("ccu_comm_create_all: no devices"); return nullptr; }
  std::vector<int> dv(n_devices);
  for (int k = 0; k < n_devices; ++k) dv[k] = devices ? devices[k] : k;
  std::string err;
  ccu::Comm* c = ccu::comm_create_all(dv, &err);
  if (!c) { fail("%s", err.c_str()); return nullptr; }
  ccu_comm* h = new ccu_comm();
  h->c = c;
  return h;
}

ccu_comm* ccu_comm_create_rank(const unsigned char id[128], int rank, int n_ranks, int device) {
  if (!id || rank < 0 || rank >= n_ranks) { fail("ccu_comm_create_rank: invalid arguments"); return nullptr; }
  std::string err;
  ccu::Comm* c = ccu::comm_create_rank(id, rank, n_ranks, device, &err);
  if (!c) { fail("%s", err.c_str()); return nullptr; }
  ccu_comm* h = new ccu_comm();
  h->c = c;
  return h;
}

void ccu_comm_destroy(ccu_comm* h) {
  if (!h) return;
  ccu::comm_destroy(h->c);
  delete h;
}

int ccu_comm_size(const ccu_comm* h) { return h ? ccu::comm_size(h->c) : 0; }
int ccu_comm_nccl_version(void) { return ccu::comm_nccl_version(); }

int ccu_comm_allreduce_block_sums(ccu_comm* h, double* const* d_part, ccu_int count, void* const* streams) {
  if (!h || !d_part) return fail("ccu_comm_allreduce_block_sums: null argument");
  const int nl = ccu::comm_local_size(h->c);
  std::vector<cudaStream_t> st(nl, nullptr);
  for (int k = 0; k < nl; ++k) st[k] = streams ? static_cast<cudaStream_t>(streams[k]) : nullptr;
  std::string err;
  if (!ccu::comm_allreduce_bits(h->c, d_part, count, st.data(), &err)) return fail("%s", err.c_str());
  g_launches += nl;
  return 0;
}

// ---------------------------------------------------------------- one tape replicated on several devices
struct ccu_multi {
  std::vector<ccu_tape*> tapes;  // one replica per device
  ccu_comm* comm = nullptr;      // null for a single device
};

// Communicators of a device set are created once per process and shared by every map on that set (ncclCommInitAll takes
// seconds; a CasADi program creates many maps): kept until the process ends, used by one evaluation at a time.
static std::mutex g_comm_mutex;
static std::vector<std::pair<std::vector<int>, ccu_comm*>> g_comm_cache;
static std::mutex g_comm_use;  // a shared communicator carries one collective at a time

static ccu_comm* shared_comm(int n_devices, const int* devices) {
  std::vector<int> key(devices, devices + n_devices);
  std::lock_guard<std::mutex> lk(g_comm_mutex);
  for (auto& e : g_comm_cache) if (e.first == key) return e.second;
  ccu_comm* c = ccu_comm_create_all(n_devices, devices);
  if (c) g_comm_cache.emplace_back(key, c);
  return c;
}

static ccu_multi* multi_finish(std::vector<ccu_tape*>& tapes, int n_devices, const int* devices) {
  ccu_multi* m = new ccu_multi();
  m->tapes = tapes;
  if (n_devices > 1) {
    m->comm = shared_comm(n_devices, devices);
    if (!m->comm) {
      const std::string e = g_err;
      for (ccu_tape* t : tapes) ccu_tape_destroy(t);
      delete m;
      g_err = "multi-device map needs NCCL: " + e;
      return nullptr;
    }
  }
  return m;
}

ccu_multi* ccu_multi_create(ccu_int n_instr, const int* op, const int* i0, const int* i1, const int* i2, const double* d,
                            ccu_int sz_w, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out, const ccu_int* nnz_out,
                            int n_devices, const int* devices) {
  if (n_devices < 1 || !devices) { fail("ccu_multi_create: no devices"); return nullptr; }
  std::vector<ccu_tape*> tapes;
  for (int k = 0; k < n_devices; ++k) {
    // (the specialised kernels of the replicas come out of the on-disk cubin cache after the first device)
    ccu_tape* t = ccu_tape_create(n_instr, op, i0, i1, i2, d, sz_w, n_in, nnz_in, n_out, nnz_out, devices[k]);
    if (!t) { for (ccu_tape* u : tapes) ccu_tape_destroy(u); return nullptr; }
    tapes.push_back(t);
  }
  return multi_finish(tapes, n_devices, devices);
}

ccu_multi* ccu_builder_finish_multi(ccu_builder* b, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out, const ccu_int* nnz_out,
                                    int n_devices, const int* devices) {
  if (n_devices < 1 || !devices) { fail("ccu_builder_finish_multi: no devices"); return nullptr; }
  std::vector<ccu_tape*> tapes;
  for (int k = 0; k < n_devices; ++k) {
    ccu_tape* t = ccu_builder_finish(b, n_in, nnz_in, n_out, nnz_out, devices[k]);
    if (!t) { for (ccu_tape* u : tapes) ccu_tape_destroy(u); return nullptr; }
    tapes.push_back(t);
  }
  return multi_finish(tapes, n_devices, devices);
}

void ccu_multi_destroy(ccu_multi* m) {
  if (!m) return;
  for (ccu_tape* t : m->tapes) ccu_tape_destroy(t);
  // (the communicator is shared: shared_comm)
  delete m;
}

int ccu_multi_size(const ccu_multi* m) { return m ? static_cast<int>(m->tapes.size()) : 0; }
ccu_tape* ccu_multi_tape(ccu_multi* m, int k) {
  return (m && k >= 0 && k < static_cast<int>(m->tapes.size())) ? m->tapes[k] : nullptr;
}

// Instances [g*N/G, (g+1)*N/G) in whole reduction blocks on device g, each shard through the chunked host pipeline of its
// own device from its own host thread; reduced outputs: block sums at global rows, one NCCL all-reduce over the devices,
// level-1 tree on device 0 (HorzRepsum / MapSum semantics, repmat.cpp:127-135, mapsum.cpp:170-184).
int ccu_multi_eval_host(ccu_multi* m, ccu_int N, const double* const* arg, double* const* res, const int* reduce_in,
                        const int* reduce_out) {
  return ccu_multi_eval_host_grouped(m, N, arg, res, reduce_in, reduce_out, nullptr, nullptr);
}

int ccu_multi_eval_host_grouped(ccu_multi* m, ccu_int N, const double* const* arg, double* const* res, const int* reduce_in,
                                const int* reduce_out, const int* in_groups, const int* out_groups) {
  if (!m || m->tapes.empty()) return fail("null multi-device map");
  const int G = static_cast<int>(m->tapes.size());
  if (G == 1) return eval_host_impl(m->tapes[0], N, arg, res, reduce_in, reduce_out, 0, N, true, in_groups, out_groups);
  if (check_eval_args(m->tapes[0], N)) return 1;
  ccu_tape* t0 = m->tapes[0];
  const size_t n_in = t0->nnz_in.size(), n_out = t0->nnz_out.size();
  const long long blocks = (N + ccu::kReduceBlock - 1) / ccu::kReduceBlock;
  std::vector<long long> off(G + 1, 0);
  for (int g = 0; g <= G; ++g) {
    const long long per = blocks / G, extra = blocks % G;
    const long long b0 = g * per + std::min<long long>(g, extra);
    off[g] = std::min<long long>(b0 * ccu::kReduceBlock, N);
  }
  if (arg && res) register_caller_buffers(t0, N, arg, res, reduce_in, reduce_out);
  bool any_red = false;
  for (size_t j = 0; j < n_out; ++j) any_red = any_red || (reduce_out && reduce_out[j] && res[j] && t0->nnz_out[j] > 0);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g) {
    th.emplace_back([&, g] {
      const long long i0 = off[g], n = off[g + 1] - off[g];
      std::vector<const double*> a(n_in, nullptr);
      std::vector<double*> r(n_out, nullptr);
      // (grouped buffers keep their base: the piece-major layout is addressed from the start of the whole batch)
      for (size_t j = 0; j < n_in; ++j)
        a[j] = arg[j] ? (((reduce_in && reduce_in[j]) || (in_groups && in_groups[j] > 1)) ? arg[j] : arg[j] + i0 * t0->nnz_in[j]) : nullptr;
      for (size_t j = 0; j < n_out; ++j)
        r[j] = res[j] ? (((reduce_out && reduce_out[j]) || (out_groups && out_groups[j] > 1)) ? res[j] : res[j] + i0 * t0->nnz_out[j]) : nullptr;
      rcs[g] = eval_host_impl(m->tapes[g], n, a.data(), r.data(), reduce_in, reduce_out, i0, N, /*finish_reduce=*/false, in_groups, out_groups);
      if (rcs[g]) errs[g] = g_err;
    });
  }
  for (auto& x : th) x.join();
  for (int g = 0; g < G; ++g) if (rcs[g]) return fail("device %d: %s", m->tapes[g]->device, errs[g].c_str());
  if (!any_red) return 0;
  for (size_t j = 0; j < n_out; ++j) {
    if (!(reduce_out && reduce_out[j] && res[j] && t0->nnz_out[j] > 0)) continue;
    const int nnz = static_cast<int>(t0->nnz_out[j]);
    std::vector<double*> bufs(G);
    std::vector<void*> streams(G);
    for (int g = 0; g < G; ++g) { bufs[g] = m->tapes[g]->d_part[j].p; streams[g] = m->tapes[g]->pipe.s[1]; }
    std::lock_guard<std::mutex> use(g_comm_use);
    if (ccu_comm_allreduce_block_sums(m->comm, bufs.data(), std::max<long long>(blocks, 1) * nnz, streams.data())) return 1;
    CCU_CUDA(cudaSetDevice(t0->device));
    HostPipe& hp = t0->pipe;
    if (hp.red.ensure(static_cast<size_t>(nnz)) || hp.red_pin.ensure(static_cast<size_t>(nnz))) return 1;
    CCU_CUDA(ccu::launch_tree(t0->d_part[j].p, blocks, nnz, hp.red.p, hp.s[1]));
    g_launches++;
    CCU_CUDA(cudaMemcpyAsync(hp.red_pin.p, hp.red.p, static_cast<size_t>(nnz) * 8, cudaMemcpyDeviceToHost, hp.s[1]));
    for (int g = 0; g < G; ++g) {
      CCU_CUDA(cudaSetDevice(m->tapes[g]->device));
      CCU_CUDA(cudaStreamSynchronize(m->tapes[g]->pipe.s[1]));
    }
    std::memcpy(res[j], hp.red_pin.p, static_cast<size_t>(nnz) * 8);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------- tape builder
struct ccu_builder {
  ccu::TapeBuilder b;
};

struct ccu_linsol {
  ccu_tape* tape = nullptr;
  ccu_int n = 0, nnz_a = 0, nrhs = 0;
  int kind = 0;  // 0 = LDL, 1 = QR
};

namespace {
ccu_int pattern_nnz(const ccu_int* sp) { return sp[2 + sp[1]]; }
bool pattern_ok(const ccu_int* sp) {
  if (!sp || sp[0] < 0 || sp[1] < 0) return false;
  const ccu_int ncol = sp[1];
  const ccu_int* ci = sp + 2;
  if (ci[0] != 0) return false;
  for (ccu_int c = 0; c < ncol; ++c) if (ci[c + 1] < ci[c]) return false;
  const ccu_int* r = sp + 2 + ncol + 1;
  for (ccu_int k = 0; k < ci[ncol]; ++k) if (r[k] < 0 || r[k] >= sp[0]) return false;
  return true;
}
bool perm_ok(const ccu_int* p, ccu_int n) {
  if (!p) return false;
  std::vector<char> seen(static_cast<size_t>(n), 0);
  for (ccu_int i = 0; i < n; ++i) {
    if (p[i] < 0 || p[i] >= n || seen[p[i]]) return false;
    seen[p[i]] = 1;
  }
  return true;
}
bool ids_ok(const ccu_builder* b, const ccu_int* v, ccu_int n) {
  for (ccu_int i = 0; i < n; ++i) if (v[i] < 0 || v[i] >= b->b.n_values()) return false;
  return true;
}
}  // namespace

ccu_builder* ccu_builder_create(void) { return new ccu_builder(); }
void ccu_builder_destroy(ccu_builder* b) { delete b; }
ccu_int ccu_builder_const(ccu_builder* b, double c) { return b ? b->b.constant(c) : -1; }
ccu_int ccu_builder_input(ccu_builder* b, ccu_int idx, ccu_int nz) {
  if (!b || idx < 0 || nz < 0) { fail("ccu_builder_input: invalid arguments"); return -1; }
  return b->b.input(idx, nz);
}
ccu_int ccu_builder_op(ccu_builder* b, int op, ccu_int x, ccu_int y) {
  if (!b || x < 0 || x >= b->b.n_values() || y >= b->b.n_values()) { fail("ccu_builder_op: invalid operand"); return -1; }
  return b->b.op(op, x, y);
}
int ccu_builder_output(ccu_builder* b, ccu_int idx, ccu_int nz, ccu_int v) {
  if (!b || idx < 0 || nz < 0 || v < 0 || v >= b->b.n_values()) return fail("ccu_builder_output: invalid arguments");
  b->b.output(idx, nz, v);
  return 0;
}

int ccu_builder_ldl(ccu_builder* b, const ccu_int* sp_a, const ccu_int* sp_lt, const ccu_int* p, const ccu_int* a,
                    ccu_int* x, ccu_int nrhs, ccu_int* zero_pivots) {
  if (!b || !pattern_ok(sp_a) || !pattern_ok(sp_lt) || sp_a[0] != sp_a[1] || sp_lt[0] != sp_a[0] || sp_lt[1] != sp_a[1] ||
      !perm_ok(p, sp_a[1]) || !a || !x || nrhs < 0)
    return fail("ccu_builder_ldl: invalid pattern, permutation or arguments");
  const ccu_int n = sp_a[1];
  if (!ids_ok(b, a, pattern_nnz(sp_a)) || !ids_ok(b, x, n * nrhs)) return fail("ccu_builder_ldl: invalid value handle");
  std::vector<ccu::TapeBuilder::V> lt, d;
  b->b.ldl(sp_a, a, sp_lt, &lt, &d, p);
  b->b.ldl_solve(x, nrhs, sp_lt, lt.data(), d.data(), p);
  if (zero_pivots) *zero_pivots = b->b.ldl_zero_pivots(d.data(), n);
  return 0;
}

int ccu_builder_qr(ccu_builder* b, const ccu_int* sp_a, const ccu_int* sp_v, const ccu_int* sp_r, const ccu_int* prinv,
                   const ccu_int* pc, const ccu_int* a, ccu_int* x, ccu_int nrhs, int tr, double eps, ccu_int* nullity) {
  if (!b || !pattern_ok(sp_a) || !pattern_ok(sp_v) || !pattern_ok(sp_r) || sp_a[0] != sp_a[1] || sp_v[1] != sp_a[1] ||
      sp_r[1] != sp_a[1] || sp_v[0] < sp_a[0] || !perm_ok(pc, sp_a[1]) || !prinv || !a || !x || nrhs < 0)
    return fail("ccu_builder_qr: invalid pattern, permutation or arguments");
  const ccu_int n = sp_a[1];
  for (ccu_int i = 0; i < sp_a[0]; ++i) if (prinv[i] < 0 || prinv[i] >= sp_v[0]) return fail("ccu_builder_qr: invalid prinv");
  if (!ids_ok(b, a, pattern_nnz(sp_a)) || !ids_ok(b, x, n * nrhs)) return fail("ccu_builder_qr: invalid value handle");
  std::vector<ccu::TapeBuilder::V> v, r, beta;
  b->b.qr(sp_a, a, sp_v, &v, sp_r, &r, &beta, prinv, pc);
  b->b.qr_solve(x, nrhs, tr != 0, sp_v, v.data(), sp_r, r.data(), beta.data(), prinv, pc);
  if (nullity) *nullity = b->b.qr_nullity(r.data(), sp_r, eps);
  return 0;
}

int ccu_builder_mtimes(ccu_builder* b, const ccu_int* x, const ccu_int* sp_x, const ccu_int* y, const ccu_int* sp_y,
                       ccu_int* z, const ccu_int* sp_z) {
  if (!b || !pattern_ok(sp_x) || !pattern_ok(sp_y) || !pattern_ok(sp_z) || sp_x[1] != sp_y[0] || sp_z[0] != sp_x[0] ||
      sp_z[1] != sp_y[1] || !x || !y || !z)
    return fail("ccu_builder_mtimes: inconsistent patterns");
  if (!ids_ok(b, x, pattern_nnz(sp_x)) || !ids_ok(b, y, pattern_nnz(sp_y)) || !ids_ok(b, z, pattern_nnz(sp_z)))
    return fail("ccu_builder_mtimes: invalid value handle");
  b->b.mtimes(x, sp_x, y, sp_y, z, sp_z);
  return 0;
}

ccu_int ccu_builder_select(ccu_builder* b, ccu_int c, ccu_int x, ccu_int y) {
  if (!b) { fail("null builder"); return -1; }
  const ccu_int n = b->b.n_values();
  if (c < 0 || c >= n || x < 0 || x >= n || y < 0 || y >= n) { fail("ccu_builder_select: invalid operand handle"); return -1; }
  return b->b.select(c, x, y);
}

ccu_int ccu_builder_export(const ccu_builder* b, int* op, int* i0, int* i1, int* i2, double* d, ccu_int cap, ccu_int* sz_w) {
  if (!b) { fail("null builder"); return -1; }
  const ccu::TapeSource s = b->b.source({}, {});
  const ccu_int n = std::min<ccu_int>(s.n_instr, cap > 0 ? cap : 0);
  if (op) std::copy(s.op, s.op + n, op);
  if (i0) std::copy(s.i0, s.i0 + n, i0);
  if (i1) std::copy(s.i1, s.i1 + n, i1);
  if (i2) std::copy(s.i2, s.i2 + n, i2);
  if (d) std::copy(s.d, s.d + n, d);
  if (sz_w) *sz_w = s.sz_w;
  return s.n_instr;
}

ccu_tape* ccu_builder_finish(ccu_builder* b, ccu_int n_in, const ccu_int* nnz_in, ccu_int n_out, const ccu_int* nnz_out,
                             int device) {
  if (!b) { fail("null builder"); return nullptr; }
  std::vector<long long> ni(nnz_in, nnz_in + (n_in > 0 ? n_in : 0)), no(nnz_out, nnz_out + (n_out > 0 ? n_out : 0));
  ccu::TapeSource s = b->b.source(ni, no);
  return ccu_tape_create(s.n_instr, s.op, s.i0, s.i1, s.i2, s.d, s.sz_w, n_in, nnz_in, n_out, nnz_out, device);
}

// ------------------------------------------------------------------------- batched linear solves (K3/K4)
static ccu_linsol* linsol_finish(ccu_builder& B, int kind, ccu_int n, ccu_int nnz_a, ccu_int nrhs, const ccu_int* x,
                                 ccu_int flag_count, int device) {
  for (ccu_int i = 0; i < n * nrhs; ++i) B.b.output(0, i, x[i]);
  B.b.output(1, 0, B.b.op(22 /* OP_NE */, flag_count, B.b.constant(0.0)));
  const ccu_int nnz_in[2] = {nnz_a, n * nrhs}, nnz_out[2] = {n * nrhs, 1};
  ccu_tape* t = ccu_builder_finish(&B, 2, nnz_in, 2, nnz_out, device);
  if (!t) return nullptr;
  ccu_linsol* ls = new ccu_linsol();
  ls->tape = t; ls->n = n; ls->nnz_a = nnz_a; ls->nrhs = nrhs; ls->kind = kind;
  return ls;
}

ccu_linsol* ccu_ldl_create(const ccu_int* sp_a, const ccu_int* sp_lt, const ccu_int* p, ccu_int nrhs, int device) {
  if (!pattern_ok(sp_a) || nrhs < 1) { fail("ccu_ldl_create: invalid arguments"); return nullptr; }
  ccu_builder B;
  const ccu_int n = sp_a[1], nnz = pattern_nnz(sp_a);
  std::vector<ccu_int> a(nnz), x(n * nrhs);
  for (ccu_int k = 0; k < nnz; ++k) a[k] = B.b.input(0, k);
  for (ccu_int i = 0; i < n * nrhs; ++i) x[i] = B.b.input(1, i);
  ccu_int flag = -1;
  if (ccu_builder_ldl(&B, sp_a, sp_lt, p, a.data(), x.data(), nrhs, &flag)) return nullptr;
  return linsol_finish(B, 0, n, nnz, nrhs, x.data(), flag, device);
}

ccu_linsol* ccu_qr_create(const ccu_int* sp_a, const ccu_int* sp_v, const ccu_int* sp_r, const ccu_int* prinv,
                          const ccu_int* pc, ccu_int nrhs, int tr, double eps, int device) {
  if (!pattern_ok(sp_a) || nrhs < 1) { fail("ccu_qr_create: invalid arguments"); return nullptr; }
  ccu_builder B;
  const ccu_int n = sp_a[1], nnz = pattern_nnz(sp_a);
  std::vector<ccu_int> a(nnz), x(n * nrhs);
  for (ccu_int k = 0; k < nnz; ++k) a[k] = B.b.input(0, k);
  for (ccu_int i = 0; i < n * nrhs; ++i) x[i] = B.b.input(1, i);
  ccu_int flag = -1;
  if (ccu_builder_qr(&B, sp_a, sp_v, sp_r, prinv, pc, a.data(), x.data(), nrhs, tr, eps, &flag)) return nullptr;
  return linsol_finish(B, 1, n, nnz, nrhs, x.data(), flag, device);
}

void ccu_linsol_destroy(ccu_linsol* ls) {
  if (!ls) return;
  ccu_tape_destroy(ls->tape);
  delete ls;
}

ccu_tape* ccu_linsol_tape(ccu_linsol* ls) { return ls ? ls->tape : nullptr; }

int ccu_linsol_solve_host(ccu_linsol* ls, ccu_int N, const double* A, const double* B, double* X, ccu_int* n_flagged) {
  if (!ls || !A || !B || !X) return fail("ccu_linsol_solve_host: null argument");
  double flag = 0;
  const double* arg[2] = {A, B};
  double* res[2] = {X, &flag};
  const int reduce_out[2] = {0, 1};
  if (eval_host_impl(ls->tape, N, arg, res, nullptr, reduce_out)) return 1;
  if (n_flagged) *n_flagged = static_cast<ccu_int>(flag);
  return 0;
}

int ccu_linsol_solve_device(ccu_linsol* ls, ccu_int N, const double* d_A, const double* d_B, double* d_X, double* d_flagged,
                            int layout, void* stream) {
  if (!ls || !d_A || !d_B || !d_X) return fail("ccu_linsol_solve_device: null argument");
  const double* arg[2] = {d_A, d_B};
  double* res[2] = {d_X, d_flagged};
  const int reduce_out[2] = {0, 1};
  return ccu_map_eval_reduce_device(ls->tape, N, arg, res, nullptr, reduce_out, layout, stream);
}

int ccu_tape_last_kernel_ms(ccu_tape* t, double* ms) {
  if (!t || !ms) return fail("null argument");
  if (!t->timed) return fail("no kernel has been launched yet");
  CCU_CUDA(cudaEventSynchronize(t->ev1));
  float f = 0;
  CCU_CUDA(cudaEventElapsedTime(&f, t->ev0, t->ev1));
  *ms = f;
  return 0;
}

int ccu_tape_last_eval_stats(const ccu_tape* t, double stats[6]) {
  if (!t || !stats) return fail("null argument");
  stats[0] = t->st_h2d_ms; stats[1] = t->st_kernel_ms; stats[2] = t->st_d2h_ms; stats[3] = t->st_stage_ms;
  stats[4] = t->st_wall_ms; stats[5] = t->st_staged_bytes;
  return 0;
}

ccu_int ccu_launch_count(void) { return g_launches.load(); }

int ccu_set_device(int device) { CCU_CUDA(cudaSetDevice(device)); return 0; }
void* ccu_malloc(ccu_int bytes) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) { fail("cudaMalloc(%lld) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
int ccu_free(void* p) { CCU_CUDA(cudaFree(p)); return 0; }
void* ccu_malloc_host(ccu_int bytes) {
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) { fail("cudaMallocHost(%lld) failed: %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
int ccu_free_host(void* p) { CCU_CUDA(cudaFreeHost(p)); return 0; }
int ccu_host_register(const void* p, ccu_int bytes) {
  if (!p || bytes <= 0) return fail("ccu_host_register: null buffer or no bytes");
  if (HostRegistry::get().pin(p, static_cast<size_t>(bytes), nullptr)) return fail("ccu_host_register: cudaHostRegister of %lld bytes at %p failed", bytes, p);
  return 0;
}
int ccu_host_unregister(const void* p) {
  if (HostRegistry::get().unpin(p)) return fail("ccu_host_unregister: %p was not registered by ccu_host_register / CCU_HOST_REGISTER", p);
  return 0;
}
int ccu_host_registered_count(void) { return static_cast<int>(HostRegistry::get().count()); }
int ccu_memcpy_h2d(void* dst, const void* src, ccu_int bytes, void* stream) {
  CCU_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}
int ccu_memcpy_d2h(void* dst, const void* src, ccu_int bytes, void* stream) {
  CCU_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return 0;
}
int ccu_stream_sync(void* stream) { CCU_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); return 0; }
int ccu_device_sync(void) { CCU_CUDA(cudaDeviceSynchronize()); return 0; }


int ccu_selftest_host_copy(long long bytes, int dst_misalign, int src_misalign, double* gb_per_s) {
  if (bytes < 0 || dst_misalign < 0 || dst_misalign > 4096 || src_misalign < 0 || src_misalign > 4096) return fail("ccu_selftest_host_copy: invalid argument");
  const size_t n = static_cast<size_t>(bytes), guard = 64;
  std::vector<unsigned char> src(n + src_misalign + 1), dst(n + dst_misalign + 2 * guard, 0xA5);
  unsigned long long x = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < n; ++i) { x = x * 6364136223846793005ull + 1442695040888963407ull; src[src_misalign + i] = static_cast<unsigned char>(x >> 56); }
  unsigned char* d = dst.data() + guard + dst_misalign;
  const auto t0 = std::chrono::steady_clock::now();
  CopyPool::get().copy(d, src.data() + src_misalign, n);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (gb_per_s) *gb_per_s = sec > 0 ? static_cast<double>(n) / sec / 1e9 : 0;
  if (n && std::memcmp(d, src.data() + src_misalign, n) != 0) return fail("ccu_selftest_host_copy: the copy differs from its source");
  for (size_t i = 0; i < guard + dst_misalign; ++i) if (dst[i] != 0xA5) return fail("ccu_selftest_host_copy: bytes before the destination were written");
  for (size_t i = 0; i < guard; ++i) if (d[n + i] != 0xA5) return fail("ccu_selftest_host_copy: bytes after the destination were written");
  return 0;
}

}  // extern "C"
