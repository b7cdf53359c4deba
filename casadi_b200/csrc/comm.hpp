// NCCL communicators for the cross-GPU reduce_out combine (comm.cu): internal C++ interface behind ccu_comm.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace ccu {

struct Comm;

bool comm_available(std::string* why);
Comm* comm_create_all(const std::vector<int>& devices, std::string* err);
Comm* comm_create_rank(const unsigned char id[128], int rank, int n_ranks, int device, std::string* err);
bool comm_unique_id(unsigned char id[128], std::string* err);
void comm_destroy(Comm* c);
int comm_local_size(const Comm* c);  // devices of this process
int comm_size(const Comm* c);        // ranks of the communicator
bool comm_allreduce_bits(Comm* c, double* const* bufs, long long count, cudaStream_t const* streams, std::string* err);
int comm_nccl_version();

}  // namespace ccu
