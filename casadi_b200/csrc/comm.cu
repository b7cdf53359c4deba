// NCCL communicators for the cross-GPU combine of reduce_out block sums (ccu_comm, include/casadi_cuda.h).
//
// The only communication of the whole path: every GPU computes the level-0 sums of its own
// 1024-instance blocks (reduce.cu) at their GLOBAL positions of a zero vector; one NCCL all-reduce over
// NVLink / NVSwitch merges the vectors; the level-1 tree is then evaluated on the merged vector, so the result has
// the bits of a single-GPU evaluation for any number of GPUs.  The supports are disjoint, hence the all-reduce runs
// on the 64-bit PATTERNS (ncclUint64, ncclSum): pattern + 0 is the pattern itself, so even -0.0 and NaN payloads
// survive, which a floating-point sum with +0.0 would not guarantee.
//
// libnccl is bound at run time (dlopen), like NVRTC: the library has no link-time dependency on it and a
// single-GPU user never loads it.  Two ways to build a communicator:
//   * ccu_comm_create_all   -- one process driving several devices (ncclCommInitAll): what CudaMap uses with
//                              CASADI_CUDA_DEVICES;
//   * ccu_comm_create_rank  -- one rank of a multi-process job (ncclCommInitRank); the 128-byte id comes from
//                              ccu_comm_unique_id on rank 0 and travels by whatever the host has (bench.py:
//                              torch.distributed broadcast).
#include "../../include/casadi_cuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "comm.hpp"

namespace ccu {

namespace {
struct Nccl {
  void* handle = nullptr;
  std::string why;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> names;
    if (const char* e = getenv("CCU_NCCL_LIB")) names.push_back(e);
    names.insert(names.end(), {"libnccl.so.2", "libnccl.so"});
    for (const auto& nm : names) {
      n.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) { n.why = std::string("libnccl not loadable: ") + dlerror(); return; }
#define CCU_SYM(f) n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.handle, "nccl" #f)); if (!n.f) { n.why = "nccl" #f " missing"; n.handle = nullptr; return; }
    CCU_SYM(GetUniqueId) CCU_SYM(CommInitRank) CCU_SYM(CommInitAll) CCU_SYM(CommDestroy) CCU_SYM(AllReduce)
    CCU_SYM(GroupStart) CCU_SYM(GroupEnd) CCU_SYM(GetErrorString) CCU_SYM(GetVersion)
#undef CCU_SYM
  });
  return n;
}
}  // namespace

bool comm_available(std::string* why) {
  Nccl& n = nccl();
  if (!n.handle && why) *why = n.why;
  return n.handle != nullptr;
}

struct Comm {
  std::vector<ncclComm_t> comms;  // one per local device
  std::vector<int> devices;
  int n_ranks = 0;
};

static std::string nccl_err(ncclResult_t r) { return nccl().GetErrorString ? nccl().GetErrorString(r) : "nccl error"; }

Comm* comm_create_all(const std::vector<int>& devices, std::string* err) {
  if (!comm_available(err)) return nullptr;
  Comm* c = new Comm();
  c->devices = devices;
  c->n_ranks = static_cast<int>(devices.size());
  c->comms.resize(devices.size());
  ncclResult_t r = nccl().CommInitAll(c->comms.data(), static_cast<int>(devices.size()), devices.data());
  if (r != ncclSuccess) { *err = "ncclCommInitAll: " + nccl_err(r); delete c; return nullptr; }
  return c;
}

Comm* comm_create_rank(const unsigned char id[128], int rank, int n_ranks, int device, std::string* err) {
  if (!comm_available(err)) return nullptr;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof uid);
  if (cudaSetDevice(device) != cudaSuccess) { *err = "cudaSetDevice failed"; return nullptr; }
  Comm* c = new Comm();
  c->devices = {device};
  c->n_ranks = n_ranks;
  c->comms.resize(1);
  ncclResult_t r = nccl().CommInitRank(&c->comms[0], n_ranks, uid, rank);
  if (r != ncclSuccess) { *err = "ncclCommInitRank: " + nccl_err(r); delete c; return nullptr; }
  return c;
}

bool comm_unique_id(unsigned char id[128], std::string* err) {
  if (!comm_available(err)) return false;
  ncclUniqueId uid;
  ncclResult_t r = nccl().GetUniqueId(&uid);
  if (r != ncclSuccess) { *err = "ncclGetUniqueId: " + nccl_err(r); return false; }
  std::memcpy(id, &uid, sizeof uid);
  return true;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  for (size_t k = 0; k < c->comms.size(); ++k) {
    cudaSetDevice(c->devices[k]);
    if (c->comms[k]) nccl().CommDestroy(c->comms[k]);
  }
  delete c;
}

int comm_local_size(const Comm* c) { return c ? static_cast<int>(c->comms.size()) : 0; }
int comm_size(const Comm* c) { return c ? c->n_ranks : 0; }

// In-place all-reduce of the 64-bit patterns of bufs[k][0..count) (one buffer per local device, on streams[k]).
bool comm_allreduce_bits(Comm* c, double* const* bufs, long long count, cudaStream_t const* streams, std::string* err) {
  if (!c || count <= 0) return true;
  Nccl& n = nccl();
  ncclResult_t r = n.GroupStart();
  for (size_t k = 0; k < c->comms.size() && r == ncclSuccess; ++k)
    r = n.AllReduce(bufs[k], bufs[k], static_cast<size_t>(count), ncclUint64, ncclSum, c->comms[k], streams[k]);
  ncclResult_t r2 = n.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { *err = "ncclAllReduce: " + nccl_err(r); return false; }
  return true;
}

int comm_nccl_version() {
  int v = 0;
  if (comm_available(nullptr) && nccl().GetVersion) nccl().GetVersion(&v);
  return v;
}

}  // namespace ccu
