// titan_comm.h -- the one collective of the path: an fp64 sum all-reduce over the ranks that share one solution
// (chromosome shards).  NCCL is bound at run time (dlopen of libnccl.so.2, preferring a copy the host
// process has already loaded), so the library itself has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

struct titan_comm {
  void *nccl = nullptr;      // ncclComm_t
  int rank = 0, world = 1, device = 0;
};

namespace titan {
// records the message titan_last_error() returns (defined next to it in titan_api.cu); returns `code`
int set_error(int code, const std::string &msg);
// in-place sum over all ranks of buf[0..n) on `st`; returns false and fills `err` on failure
bool comm_allreduce_sum_f64(titan_comm *c, double *buf, size_t n, cudaStream_t st, std::string &err);
}  // namespace titan
