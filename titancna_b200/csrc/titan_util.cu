// titan_util.cu -- measurement helpers of the C ABI: the FP64 (DFMA) peak of the device the E-step runs on.
//
// SURVEY.md 8(d) / BASELINE.md 4 ask for the FP64 roofline denominator to be MEASURED on the box (MEASURED_PEAKS.json
// holds only the HBM copy bandwidth and the bf16 tensor figure).  bench.py calls this once, outside the timed region.
#include "../../include/titancna_b200.h"

#include <cuda_runtime.h>

#include <algorithm>

namespace {

constexpr int kChains = 8;      // independent accumulators per thread: covers the DFMA latency at 8+ warps per scheduler
constexpr int kInner = 64;      // unrolled DFMAs per chain per outer iteration

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, double x0, double m, int outer) {
  double acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc[c] = x0 + (double)(threadIdx.x + c);
  for (int it = 0; it < outer; ++it) {
#pragma unroll
    for (int i = 0; i < kInner; ++i)
#pragma unroll
      for (int c = 0; c < kChains; ++c) acc[c] = fma(acc[c], m, x0);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += acc[c];
  if (s == 12345.6789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

}  // namespace

// tflops_out: best-of-`reps` DFMA throughput (2 flop per DFMA) with every SM busy; ms_out: duration of that launch.
extern "C" int titan_measure_fp64_peak(double *tflops_out, double *ms_out) {
  if (!tflops_out) return TITAN_ERR_ARG;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return TITAN_ERR_NO_DEVICE;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return TITAN_ERR_NO_DEVICE;
  const int blocks = sms * 8, threads = 256, outer = 512, reps = 5;
  double *sink = nullptr;
  if (cudaMalloc(&sink, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return TITAN_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {           // launch 0 is the warm-up
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(sink, 1.0, 0.999999, outer);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return TITAN_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const double flop = 2.0 * (double)blocks * threads * kChains * kInner * (double)outer;
  *tflops_out = flop / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return TITAN_OK;
}

// ------------------------------------------------------------------------------------------------------------
// NCCL binding (run-time): chromosome-sharded solutions all-reduce their sufficient statistics once per E-step
// (reference: the sum over foreach workers in R/hmmClonal.R:203-234)
// ------------------------------------------------------------------------------------------------------------
#include <dlfcn.h>

#include <mutex>
#include <string>

#include "titan_comm.h"

namespace {

struct NcclId { char internal[128]; };     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128 in every 2.x release)
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
  void *handle = nullptr;
  int (*GetVersion)(int *) = nullptr;
  int (*GetUniqueId)(NcclId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string why;
};

NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);      // the copy the host process already uses
    for (const char *n : names)
      if (!api.handle) api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char *s) { return dlsym(api.handle, s); };
    api.GetVersion = reinterpret_cast<int (*)(int *)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId *)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void **, int, NcclId, int)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(void *)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<const char *(*)(int)>(sym("ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
      api.why = "libnccl.so.2 lacks the expected entry points";
      api.handle = nullptr;
    }
  });
  return api;
}

int comm_fail(int code, const std::string &msg) { return titan::set_error(code, msg); }
std::string nccl_msg(NcclApi &api, const char *what, int rc) {
  return std::string(what) + " failed: " + (api.GetErrorString ? api.GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")";
}

}  // namespace

extern "C" int titan_comm_unique_id(char *id_out) {
  if (!id_out) return comm_fail(TITAN_ERR_ARG, "titan_comm_unique_id: NULL argument");
  NcclApi &api = nccl_api();
  if (!api.handle) return comm_fail(TITAN_ERR_UNSUPPORTED, api.why);
  NcclId id;
  const int rc = api.GetUniqueId(&id);
  if (rc != 0) return comm_fail(TITAN_ERR_CUDA, nccl_msg(api, "ncclGetUniqueId", rc));
  for (int i = 0; i < TITAN_COMM_ID_BYTES; ++i) id_out[i] = id.internal[i];
  return TITAN_OK;
}

extern "C" int titan_comm_create(const char *id, int rank, int world, titan_comm **out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return comm_fail(TITAN_ERR_ARG, "titan_comm_create: bad argument");
  NcclApi &api = nccl_api();
  if (!api.handle) return comm_fail(TITAN_ERR_UNSUPPORTED, api.why);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return comm_fail(TITAN_ERR_NO_DEVICE, "titan_comm_create: no CUDA device");
  NcclId nid;
  for (int i = 0; i < TITAN_COMM_ID_BYTES; ++i) nid.internal[i] = id[i];
  void *comm = nullptr;
  const int rc = api.CommInitRank(&comm, world, nid, rank);
  if (rc != 0) return comm_fail(TITAN_ERR_CUDA, nccl_msg(api, "ncclCommInitRank", rc));
  titan_comm *c = new titan_comm;
  c->nccl = comm; c->rank = rank; c->world = world; c->device = dev;
  // NCCL connects its channels lazily inside the first collective (hundreds of milliseconds): do that here, once,
  // so that the first E-step of a sharded solution costs what every other one costs
  double *warm = nullptr;
  if (cudaMalloc(&warm, sizeof(double) * 2048) == cudaSuccess) {
    cudaMemset(warm, 0, sizeof(double) * 2048);
    const int rc2 = api.AllReduce(warm, warm, 2048, kNcclFloat64, kNcclSum, comm, (cudaStream_t)0);
    const cudaError_t ce = cudaStreamSynchronize((cudaStream_t)0);
    cudaFree(warm);
    if (rc2 != 0 || ce != cudaSuccess) {
      api.CommDestroy(comm);
      delete c;
      return comm_fail(TITAN_ERR_CUDA, rc2 != 0 ? nccl_msg(api, "ncclAllReduce (warm-up)", rc2) : std::string("warm-up all-reduce failed: ") + cudaGetErrorString(ce));
    }
  }
  *out = c;
  return TITAN_OK;
}

extern "C" void titan_comm_destroy(titan_comm *c) {
  if (!c) return;
  NcclApi &api = nccl_api();
  if (api.handle && c->nccl) api.CommDestroy(c->nccl);
  delete c;
}

extern "C" int titan_comm_info(const titan_comm *c, int *rank, int *world, int *nccl_version) {
  if (!c) return comm_fail(TITAN_ERR_ARG, "titan_comm_info: NULL argument");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (nccl_version) {
    *nccl_version = 0;
    NcclApi &api = nccl_api();
    if (api.handle && api.GetVersion) api.GetVersion(nccl_version);
  }
  return TITAN_OK;
}

namespace titan {
bool comm_allreduce_sum_f64(titan_comm *c, double *buf, size_t n, cudaStream_t st, std::string &err) {
  NcclApi &api = nccl_api();
  if (!api.handle || !c || !c->nccl) { err = api.handle ? "no communicator" : api.why; return false; }
  const int rc = api.AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, c->nccl, st);
  if (rc != 0) { err = nccl_msg(api, "ncclAllReduce", rc); return false; }
  return true;
}
}  // namespace titan
