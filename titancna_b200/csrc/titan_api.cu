// titan_api.cu -- C ABI (include/titancna_b200.h) and host orchestration of the sm_100a kernels.
//
// There is deliberately no CPU implementation of the hot path in this file: every compute entry
// fails with TITAN_ERR_NO_DEVICE when no CUDA device is usable.
#include "../../include/titancna_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "titan_comm.h"
#include "titan_fb2.h"
#include "titan_host.hpp"
#include "titan_kernels.cuh"

using namespace titan;
namespace th = titan_host;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

struct CudaFail {
  int code;
};

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      int code_ = (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? TITAN_ERR_NO_DEVICE \
                                                                                 : TITAN_ERR_CUDA;  \
      fail(code_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);      \
      throw CudaFail{code_};                                                                        \
    }                                                                                               \
  } while (0)

extern "C" const char *titan_last_error(void) { return g_err.c_str(); }
int titan::set_error(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
extern "C" const char *titan_version(void) { return "titancna_b200 0.1 (sm_100a)"; }

extern "C" int titan_device_count(int *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (count) *count = 0;
    return fail(TITAN_ERR_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(e));
  }
  if (count) *count = n;
  return n > 0 ? TITAN_OK : fail(TITAN_ERR_NO_DEVICE, "no CUDA device present (this library has no CPU fallback)");
}

extern "C" int titan_get_device(int *device) {
  if (!device) return fail(TITAN_ERR_ARG, "titan_get_device: device is NULL");
  cudaError_t e = cudaGetDevice(device);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(TITAN_ERR_NO_DEVICE, "no usable CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  }
  return TITAN_OK;
}

extern "C" int titan_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(TITAN_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  return TITAN_OK;
}

// ---------------------------------------------------------------------------------------------
// solver
// ---------------------------------------------------------------------------------------------
// Device buffers come from the stream-ordered allocator with a pool that never shrinks on its own: a sweep
// creates and destroys a solver (hundreds of MB) per solution, and cudaMalloc / cudaFree of such blocks
// sporadically cost 50-500 ms each (unmap / remap), several times the E-step.  Ordering contract: every API
// entry leaves its streams synchronised and ~titan_solver synchronises before its members are released, so
// a block is never returned to the pool while work on it is in flight.
// The pool is the library's OWN (one per device, created on first use): the device's default pool, which a host
// application may share, keeps its release threshold.
static cudaMemPool_t titan_pool(int dev) {
  static std::mutex mu;
  static std::unordered_map<int, cudaMemPool_t> pools;
  std::lock_guard<std::mutex> lock(mu);
  auto it = pools.find(dev);
  if (it != pools.end()) return it->second;
  cudaMemPoolProps props{};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  cudaMemPool_t pool = nullptr;
  if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
    cudaGetLastError();
    pool = nullptr;                                   // fall back to the default pool, threshold untouched
  } else {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  pools[dev] = pool;
  return pool;
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    cudaMemPool_t pool = titan_pool(dev);
    if (pool) CK(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&p), count * sizeof(T), pool, cudaStreamPerThread));
    else CK(cudaMallocAsync(reinterpret_cast<void **>(&p), count * sizeof(T), cudaStreamPerThread));
    CK(cudaStreamSynchronize(cudaStreamPerThread));     // usable from any stream from here on
    n = count;
  }
  void release() {
    if (p) cudaFreeAsync(p, cudaStreamPerThread);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

// Pinned host staging buffers are a few KB per solver, but cudaMallocHost / cudaFreeHost sporadically take 80-500 ms
// (page pinning / unpinning; measured in a 15-solution sweep, TITAN_TRACE=1): blocks are recycled through a
// process-wide cache of power-of-two size classes and only given back at process exit.
struct PinnedCache {
  std::mutex mu;
  std::unordered_map<size_t, std::vector<void *>> free_blocks;      // bytes (power of two >= 4 KB) -> blocks
  static size_t size_class(size_t bytes) {
    size_t c = 4096;
    while (c < bytes) c <<= 1;
    return c;
  }
  void *get(size_t bytes, size_t *cls_out) {
    const size_t cls = size_class(bytes);
    *cls_out = cls;
    {
      std::lock_guard<std::mutex> lock(mu);
      auto &v = free_blocks[cls];
      if (!v.empty()) {
        void *p = v.back();
        v.pop_back();
        return p;
      }
    }
    void *p = nullptr;
    CK(cudaMallocHost(&p, cls));
    return p;
  }
  void put(void *p, size_t cls) {
    std::lock_guard<std::mutex> lock(mu);
    free_blocks[cls].push_back(p);
  }
};
static PinnedCache &pinned_cache() {
  static PinnedCache *c = new PinnedCache();      // never destroyed: the blocks must outlive every solver
  return *c;
}

template <class T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0, cls = 0;
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    p = static_cast<T *>(pinned_cache().get(count * sizeof(T), &cls));
    n = count;
  }
  void release() {
    if (p) pinned_cache().put(p, cls);
    p = nullptr;
    n = 0;
  }
  ~PinBuf() { release(); }
};

// TITAN_TRACE=1: wall-clock phases of solver creation and destruction on stderr
struct PhaseTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  PhaseTrace() : on(getenv("TITAN_TRACE") && atoi(getenv("TITAN_TRACE"))), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[titan trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

static const int kMaxZ = 8;
static const int kRoundSlots = 4096;

// Host-exact, parameter-independent transition terms of one data set (positions + txnExpLen + zStrength).  A sweep
// creates one solver per (ploidy, numClusters) on the SAME positions, and hashing a million gap lengths costs more
// than everything else in solver creation (27 of ~45 ms): the tables of the most recent data sets are kept and
// shared, keyed on a 64-bit hash of the position column.
struct GapTables {
  std::vector<double> rhoG_h, rhoZ_h;   // [N] host-exact transition terms between t-1 and t
  std::vector<double> pairG, pairZ;     // the distinct (rhoG, rhoZ) pairs of this data set
  std::vector<int> pair_id;             // [N] which pair sits between SNP t-1 and t
};
struct GapKey {
  int64_t N = 0;
  int nChr = 0;
  double txn = 0, zs = 0;
  uint64_t hpos = 0, hchr = 0;
  bool operator==(const GapKey &o) const {
    return N == o.N && nChr == o.nChr && txn == o.txn && zs == o.zs && hpos == o.hpos && hchr == o.hchr;
  }
};
static uint64_t hash_words(const void *p, size_t n_words, uint64_t h) {
  const uint64_t *w = static_cast<const uint64_t *>(p);
  for (size_t i = 0; i < n_words; ++i) {
    h ^= w[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
  }
  return h;
}
struct GapCache {
  std::mutex mu;
  std::vector<std::pair<GapKey, std::shared_ptr<const GapTables>>> recent;     // newest last, at most 4
  std::shared_ptr<const GapTables> find(const GapKey &k) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : recent)
      if (e.first == k) return e.second;
    return nullptr;
  }
  void put(const GapKey &k, std::shared_ptr<const GapTables> v) {
    std::lock_guard<std::mutex> lock(mu);
    if (recent.size() >= 4) recent.erase(recent.begin());
    recent.emplace_back(k, std::move(v));
  }
};
static GapCache &gap_cache() {
  static GapCache *c = new GapCache();
  return *c;
}

// Device-resident decode tables of one (positions, transition settings, state layout): the exact log-transition
// table of every distinct (rhoG, rhoZ) pair and the table id of every SNP.  Parameter-independent, 32 K bytes per pair
// (50-80 MB for a genome) and ~40-100 ms of host libm to build, so the solutions of a sweep that share a cluster count
// (same K: the three ploidy initialisations) share one copy.
struct VitTables {
  int device = 0;
  DevBuf<double> ltab;             // [pairs][K][4]
  DevBuf<int> txn_id;              // [N]
  bool monotone = false;           // every row: logA(class 2) >= logA(class 0) and logA(class 3) >= logA(class 1)
  ~VitTables() {
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
      cudaSetDevice(device);
      ltab.release();
      txn_id.release();
      cudaSetDevice(cur);
    }
  }
};
struct VitKey {
  GapKey gk;
  int K = 0, U = 0, device = 0;
  std::vector<unsigned char> het;
  bool operator==(const VitKey &o) const { return gk == o.gk && K == o.K && U == o.U && device == o.device && het == o.het; }
};
struct VitCache {
  std::mutex mu;
  std::vector<std::pair<VitKey, std::shared_ptr<const VitTables>>> recent;     // newest last, at most 6
  std::shared_ptr<const VitTables> find(const VitKey &k) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : recent)
      if (e.first == k) return e.second;
    return nullptr;
  }
  void put(const VitKey &k, std::shared_ptr<const VitTables> v) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : recent)
      if (e.first == k) { e.second = std::move(v); return; }
    if (recent.size() >= 6) recent.erase(recent.begin());
    recent.emplace_back(k, std::move(v));
  }
  void clear() {
    std::lock_guard<std::mutex> lock(mu);
    recent.clear();
  }
};
static VitCache &vit_cache() {
  static VitCache *c = new VitCache();
  return *c;
}

// The structured transition coefficients of every SNP (128 bytes each, derived on the device from the host-exact
// rhoG / rhoZ): parameter-independent like the decode tables and shared the same way.
struct TxnTable {
  int device = 0;
  DevBuf<TxnRec> rec;              // [N + 2 kFbPadRows]
  ~TxnTable() {
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
      cudaSetDevice(device);
      rec.release();
      cudaSetDevice(cur);
    }
  }
};
struct TxnKey {
  GapKey gk;
  int K = 0, U = 0, Z = 0, h = 0, device = 0;
  bool operator==(const TxnKey &o) const {
    return gk == o.gk && K == o.K && U == o.U && Z == o.Z && h == o.h && device == o.device;
  }
};
struct TxnCache {
  std::mutex mu;
  std::vector<std::pair<TxnKey, std::shared_ptr<const TxnTable>>> recent;      // newest last, at most 6
  std::shared_ptr<const TxnTable> find(const TxnKey &k) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : recent)
      if (e.first == k) return e.second;
    return nullptr;
  }
  void put(const TxnKey &k, std::shared_ptr<const TxnTable> v) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : recent)
      if (e.first == k) { e.second = std::move(v); return; }
    if (recent.size() >= 6) recent.erase(recent.begin());
    recent.emplace_back(k, std::move(v));
  }
  void clear() {
    std::lock_guard<std::mutex> lock(mu);
    recent.clear();
  }
};
static TxnCache &txn_cache() {
  static TxnCache *c = new TxnCache();
  return *c;
}

struct titan_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t stream2 = nullptr;   // speculative final pass (created on first use)
  cudaEvent_t ev_spec[2] = {nullptr, nullptr};
  int spec_final = 1;               // TITAN_SPEC_FINAL
  DevBuf<unsigned int> ver;         // [2][nChunks] run counters / counters seen by the last final pass
  int64_t N = 0;
  int nChr = 0, U = 0, Z = 0, K = 0, h = 0;
  bool legacy_py = false;          // emissions come from a materialised py matrix
  bool gauss = false;              // alleleEmissionModel == "Gaussian": bivariate normal on (ref / tumDepth, logR)
  std::vector<int64_t> chr_start;
  std::vector<unsigned char> het;
  std::shared_ptr<const GapTables> gaps;   // host-exact transition terms of this data set (shared between solvers)

  // configuration
  int chunk_len = 1024, warm = 192, max_rounds = 0;   // max_rounds 0: bounded by the chunk count (titan_solver_configure)
  int engine = 2;                  // titan_fb2: materialised emissions, posterior-weighted boundary test, probe-accelerated rounds
  int Kp = 0;                      // row stride of the per-SNP K-vectors in HBM: K states + the emission shift, padded to 16 bytes
  double tol = 1e-8;               // posterior-weighted boundary tolerance of the fb2 engine
  int vit_mode = -1;               // 4: warp-reduction scan, 3: cluster-sliced scan, 2: register-sliced dense scan, 0: shared-memory dense scan, -1: by K (TITAN_VIT_MODE)

  // device data
  DevBuf<SnpRec> snp;
  std::shared_ptr<const TxnTable> txn;     // shared between solvers of the same positions and state layout (TxnCache)
  DevBuf<double> py;               // legacy mode only
  DevBuf<Chunk> chunks;
  std::vector<Chunk> chunks_h;
  int nChunks = 0, maxChunksPerChr = 1;
  DevBuf<double> alpha, startf, endf, startb, endb, chunk_ll, part, out, rhoG0, rhoZ0, rhoG, rhoZ, loggamma;
  int probes = 1;                  // fb2 engine: probe-accelerated rounds (TITAN_PROBES)
  int probe_window = 8;            // from the second cycle on, chunks this close downstream of a dirty one are probed too (TITAN_PROBE_WINDOW)
  int probe_window1 = 0;           // the same for the first cycle (TITAN_PROBE_WINDOW1)
  int plain_after = 0;             // > 0: cycles after this one re-run dirty chunks from the neighbour's vector, no probes (TITAN_PLAIN_AFTER)
  DevBuf<unsigned char> pflags;    // [4][nChunks] dirtyF, dirtyB, solvedF, solvedB
  DevBuf<int> pints;               // [2][nChunks][2] probed states, [2][nChunks][3] probe exponents
  DevBuf<double> pvec;             // [2][nChunks][3][K] probe results, [2][nChunks][K] predicted boundary vectors
  DevBuf<double> subb;             // [nChunks][3][K] beta~ at the quarter boundaries of every chunk
  DevBuf<double> Bm, extf, extb;   // fb2 engine: scaled emissions [N][Kp]; one-step-extended boundary vectors [2][nChunks][K]
  // per-SNP arrays carry kFbPadRows rows of slack on both sides (TMA copies whole tiles)
  SnpRec *snp_p() const { return snp.p + kFbPadRows; }
  const TxnRec *txn_p() const { return txn->rec.p + kFbPadRows; }
  double *alpha_p() const { return alpha.p + (size_t)kFbPadRows * Kp; }
  double *Bm_p() const { return Bm.p + (size_t)kFbPadRows * Kp; }
  DevBuf<int> chr_chunk0;          // [nChr+1]
  int round_batch = 4;             // verification rounds queued per host synchronisation
  DevBuf<unsigned int> rerun;      // [2][kRoundSlots]
  DevBuf<double> model;            // packed ModelConsts arrays
  double model_l0 = 0, model_rho2 = 0;   // Gaussian allele model: scalars of the last upload_model
  DevBuf<unsigned char> het_d;
  PinBuf<double> model_h, out_h;
  PinBuf<unsigned int> rerun_h;

  // Viterbi (lazy)
  DevBuf<int> chr_start32, path, tile_first;
  std::shared_ptr<const VitTables> vt;   // log-transition tables + table ids (shared between solvers, VitCache)
  VitKey vit_key;
  int bt_rows = 0, nTiles = 0;
  DevBuf<unsigned char> psi, bt_comp, bt_entry;
  bool vit_ready = false;
  bool vit_monotone = false;        // every table row has logA(class 2) >= logA(class 0) and logA(class 3) >= logA(class 1)
  std::thread vit_builder;          // builds the Viterbi log-transition tables on the host while the EM runs and puts them
                                    // (and the rest of the decode's device state) on the device
  int vit_err = TITAN_OK;           // outcome of that thread, reported by the first decode
  std::string vit_errmsg;
  std::vector<double> vit_tab;      // [nDistinct][K][4]
  bool state_valid = false;        // alpha + verified backward boundaries of the last E-step are still on the device

  // chromosome-sharded solution (titan_solver_set_comm): this solver holds chromosomes chr_global[] of nChrTotal
  titan_comm *comm = nullptr;
  int nChrTotal = 0;
  std::vector<int> chr_global;
  DevBuf<int> chr_global_d;
  DevBuf<double> gvec;             // [6K | nChrTotal loglik | U x nChrTotal | Z x nChrTotal], all-reduced in place
  int nChrOut() const { return comm ? nChrTotal : nChr; }

  cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  titan_timing tm{};

  ~titan_solver() {
    PhaseTrace tr;
    if (vit_builder.joinable()) vit_builder.join();
    tr.mark("destroy: join table builder");
    if (stream) cudaStreamSynchronize(stream);          // members (device buffers) go back to the pool after this body
    if (stream2) cudaStreamSynchronize(stream2);
    tr.mark("destroy: stream sync");
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
    for (auto &e : ev_spec)
      if (e) cudaEventDestroy(e);
    if (stream2) cudaStreamDestroy(stream2);
    if (own_stream && stream) cudaStreamDestroy(stream);
    tr.mark("destroy: events + streams");
    model_h.release(); out_h.release(); rerun_h.release();
    tr.mark("destroy: pinned host buffers");
  }
};

static void build_viterbi_tables(titan_solver *s);
static void viterbi_device_setup(titan_solver *s);

static void build_chunks(titan_solver *s) {
  s->chunks_h.clear();
  s->maxChunksPerChr = 1;
  for (int c = 0; c < s->nChr; ++c) {
    int cs = (int)s->chr_start[c], ce = (int)s->chr_start[c + 1];
    if (ce <= cs) continue;
    int T = ce - cs;
    int L = s->chunk_len > 0 ? s->chunk_len : T;
    int n = (T + L - 1) / L;
    // even split so that no chunk is a short straggler
    int base = T / n, extra = T % n, t = cs;
    for (int i = 0; i < n; ++i) {
      int len = base + (i < extra ? 1 : 0);
      Chunk ch{t, t + len, cs, ce, c, 0};
      s->chunks_h.push_back(ch);
      t += len;
    }
    s->maxChunksPerChr = std::max(s->maxChunksPerChr, n);
  }
  s->nChunks = (int)s->chunks_h.size();
  s->chunks.alloc(s->nChunks);
  CK(cudaMemcpy(s->chunks.p, s->chunks_h.data(), sizeof(Chunk) * s->nChunks, cudaMemcpyHostToDevice));
  const size_t K = s->K, nc = s->nChunks;
  s->startf.alloc(nc * K);
  s->startb.alloc(nc * K);
  s->endf.alloc(2 * nc * K);
  s->endb.alloc(2 * nc * K);
  s->extf.alloc(2 * nc * K);
  s->extb.alloc(2 * nc * K);
  s->subb.alloc(3 * nc * K);
  s->pflags.alloc(6 * nc);
  s->ver.alloc(2 * nc);
  s->pints.alloc((4 * kProbes - 2) * nc);
  s->pvec.alloc((2 * kProbes + 2) * nc * K);
  CK(cudaMemset(s->pflags.p, 0, 6 * nc));
  // a chunk without a neighbour never writes its extended vector; the neighbour-less tests never read it either,
  // but keep the buffers defined
  CK(cudaMemset(s->extf.p, 0, sizeof(double) * 2 * nc * K));
  CK(cudaMemset(s->extb.p, 0, sizeof(double) * 2 * nc * K));
  s->chunk_ll.alloc(nc);
  s->part.alloc(nc * 6 * K);
  s->tm.n_chunks = s->nChunks;
  std::vector<int> cc0(s->nChr + 1, 0);
  for (const Chunk &ch : s->chunks_h) cc0[ch.chr + 1]++;
  for (int c = 0; c < s->nChr; ++c) cc0[c + 1] += cc0[c];
  s->chr_chunk0.alloc(s->nChr + 1);
  CK(cudaMemcpy(s->chr_chunk0.p, cc0.data(), sizeof(int) * (s->nChr + 1), cudaMemcpyHostToDevice));
}

struct SolverInit {
  int64_t N;
  int nChr;
  const int64_t *chr_start;
  const double *posn, *ref, *depth, *logR, *py;
  double txnExpLen, zStrength;
  int U, Z;
  const double *ZS;
  int nZS;
  int alleleModel = 0;
};


static int solver_create_impl(const SolverInit &in, titan_solver **out) {
  PhaseTrace trace;
  if (!out) return fail(TITAN_ERR_ARG, "titan_solver_create: out is NULL");
  *out = nullptr;
  if (in.N <= 0 || in.nChr <= 0 || !in.chr_start || !in.posn)
    return fail(TITAN_ERR_ARG, "titan_solver_create: empty data (N=%lld, nChr=%d)", (long long)in.N, in.nChr);
  if (in.N > 2000000000LL) return fail(TITAN_ERR_UNSUPPORTED, "N=%lld exceeds the 32-bit SNP index range", (long long)in.N);
  if (in.Z < 1 || in.Z > kMaxZ)
    return fail(TITAN_ERR_UNSUPPORTED, "numClusters=%d outside the supported range 1..%d", in.Z, kMaxZ);
  if (in.U < 1 || in.U > 64)
    return fail(TITAN_ERR_UNSUPPORTED, "%d genotype states: the warp-per-chunk kernels cover up to 64 (two per lane)", in.U);
  if (!in.ZS || in.nZS < in.U) return fail(TITAN_ERR_ARG, "zygosityKey must have %d entries", in.U);
  if (in.alleleModel != TITAN_ALLELE_BINOMIAL && in.alleleModel != TITAN_ALLELE_GAUSSIAN)
    return fail(TITAN_ERR_ARG, "loadDefaultParameters: alleleEmissionModel must be either \"binomial\" or \"Gaussian\".");
  if (!(in.txnExpLen > 0) || !(in.zStrength > 0))
    return fail(TITAN_ERR_ARG, "txnExpLen / zStrength must be positive scalars (length-0 arguments are rejected, "
                               "cf. R/utils.R:1514-1516)");
  if (in.chr_start[0] != 0 || in.chr_start[in.nChr] != in.N)
    return fail(TITAN_ERR_ARG, "chr_start must run from 0 to N");
  for (int c = 0; c < in.nChr; ++c)
    if (in.chr_start[c + 1] < in.chr_start[c]) return fail(TITAN_ERR_ARG, "chr_start must be non-decreasing");
  for (int a = 0; a < in.U; ++a)
    for (int b = a + 1; b < in.U; ++b)
      if (in.ZS[a] == in.ZS[b] && in.ZS[a] != -1)
        return fail(TITAN_ERR_UNSUPPORTED, "zygosityKey entries %d and %d are equal (%g): the structured transition "
                                           "needs distinct zygosity classes", a, b, in.ZS[a]);
  {
    int nhet = 0;
    for (int a = 0; a < in.U; ++a) nhet += (in.ZS[a] == -1);
    if (nhet > 1)
      return fail(TITAN_ERR_UNSUPPORTED, "%d zygosityKey entries equal -1: the reference puts them in one zygosity class "
                                         "(src/fwd_backC_clonalCN.c:262), the structured transition covers a single diploid-HET state", nhet);
  }
  int ndev = 0;
  int rc = titan_device_count(&ndev);
  if (rc != TITAN_OK) return rc;

  std::unique_ptr<titan_solver> S(new titan_solver());
  titan_solver *s = S.get();
  // Round-0 warm-up length: more clonal clusters means more nearly indistinguishable cluster pairs and a slower
  // forgetting of the start vector; a longer flat warm-up then saves whole verification cycles (measured at 1 M / 2 M
  // SNPs: Z = 4: 512 beats 192 by 15-20 %, Z = 5: 1024 by 15-25 %; Z <= 3: 192 is best).  titan_solver_configure overrides.
  s->warm = in.Z >= 5 ? 1024 : in.Z == 4 ? 512 : 192;
  // Chunk length: 1024 SNPs once the genome fills the GPU; below ~300 k SNPs every pass is latency-bound whatever the
  // chunk count, so shorter chunks shorten every pass (measured: 50 k-SNP exome 1.25 -> 0.58 ms per E-step at 128,
  // 200 k-SNP haplotype bins 1.40 -> 0.88 ms at 512; 1 M SNPs: 512 and 1024 equal).
  {
    // aim for >= 300 chunks; with four or more clusters shorter chunks cost extra verification cycles (2 M SNPs, Z = 5:
    // 512 needs 6-11 cycles where 1024 needs 4-6), so the length comes down only for half as many SNPs
    const int64_t target = in.N / (in.Z >= 4 ? 150 : 300);
    int len = 1024;
    while (len > 128 && len > target) len >>= 1;
    s->chunk_len = len;
    s->warm = std::min(s->warm, std::max(192, len));
  }
  if (const char *e = getenv("TITAN_VIT_MODE")) s->vit_mode = atoi(e);
  if (const char *e = getenv("TITAN_ROUND_BATCH")) s->round_batch = std::max(1, atoi(e));
  if (const char *e = getenv("TITAN_PROBES")) s->probes = atoi(e);
  if (const char *e = getenv("TITAN_SPEC_FINAL")) s->spec_final = atoi(e);
  if (const char *e = getenv("TITAN_PROBE_WINDOW")) s->probe_window = std::max(0, atoi(e));
  if (const char *e = getenv("TITAN_PROBE_WINDOW1")) s->probe_window1 = std::max(0, atoi(e));
  if (const char *e = getenv("TITAN_PLAIN_AFTER")) s->plain_after = std::max(0, atoi(e));
  try {
    CK(cudaGetDevice(&s->device));
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
    for (auto &e : s->ev) CK(cudaEventCreate(&e));
    trace.mark("streams + events");
    s->N = in.N; s->nChr = in.nChr; s->U = in.U; s->Z = in.Z; s->K = in.U * in.Z;
    s->Kp = (s->K + 2) & ~1;           // K states + the shift m_t, rows a multiple of 16 bytes
    s->chr_start.assign(in.chr_start, in.chr_start + in.nChr + 1);
    s->het.resize(in.U);
    s->h = 0;
    for (int u = 0; u < in.U; ++u) { s->het[u] = (in.ZS[u] == -1); s->h += s->het[u]; }
    s->legacy_py = in.py != nullptr;
    s->vit_key.K = s->K; s->vit_key.U = s->U; s->vit_key.device = s->device; s->vit_key.het = s->het;
    s->gauss = !s->legacy_py && in.alleleModel == TITAN_ALLELE_GAUSSIAN;
    const bool gauss = s->gauss;
    const int64_t N = in.N;
    const int K = s->K;

    // per-SNP records (lgamma bracket from a host table): built on worker threads while this thread hashes the gaps
    std::vector<SnpRec> rec;
    std::atomic<int64_t> bad_snp{-1};
    struct Joiner {
      std::thread t;
      void join() { if (t.joinable()) t.join(); }
      ~Joiner() { join(); }
    } rec_builder;
    if (!s->legacy_py && in.ref && in.depth && in.logR) {
      rec.resize((size_t)N);
      rec_builder.t = std::thread([&rec, &bad_snp, &in, N, gauss]() {
        double dmax = 0;
        for (int64_t t = 0; t < N; ++t) dmax = std::max(dmax, in.depth[t]);
        th::LgammaTable tab;
        if (!gauss && dmax >= 0 && dmax < 1e7) tab.ensure((int64_t)dmax);       // shared read-only table; larger depths use lgamma directly
        const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (unsigned w = 0; w < nt; ++w)
          pool.emplace_back([&, w]() {
            th::LgammaTable own;                                        // only touched for depths beyond the shared table
            const int64_t t0 = N * w / nt, t1 = N * (w + 1) / nt;
            for (int64_t t = t0; t < t1; ++t) {
              const double x = in.ref[t], d = in.depth[t];
              if (!(x == x) || !(d == d) || !(in.logR[t] == in.logR[t])) {
                int64_t expect = -1;
                bad_snp.compare_exchange_strong(expect, t);
                return;
              }
              if (gauss) {      // x1 = ref / tumDepth (R/hmmClonal.R:153); a NaN ratio makes the reference's sums NA
                const double r = x / d;
                if (!(r == r)) {
                  int64_t expect = -1;
                  bad_snp.compare_exchange_strong(expect, t);
                  return;
                }
                rec[t].x = r;
                rec[t].nx = r * r;
                rec[t].lr = in.logR[t];
                rec[t].lb = r * in.logR[t];
                continue;
              }
              rec[t].x = x;
              rec[t].nx = d - x;
              rec[t].lr = in.logR[t];
              rec[t].lb = (d >= 0 && d < (double)tab.v.size()) ? th::binom_log_norm(x, d, tab) : th::binom_log_norm(x, d, own);
            }
          });
        for (auto &th_ : pool) th_.join();
      });
    }

    // host-exact transition terms (parameter independent).  exp() is evaluated once per DISTINCT gap
    // length (an open-addressing table keyed on the gap's bit pattern), then gaps that round to the
    // same (rhoG, rhoZ) pair are merged: the Viterbi tables are built per distinct pair.
    GapKey gkey;
    gkey.N = N; gkey.nChr = in.nChr; gkey.txn = in.txnExpLen; gkey.zs = in.zStrength;
    gkey.hpos = hash_words(in.posn, (size_t)N, 1);
    gkey.hchr = hash_words(in.chr_start, (size_t)in.nChr + 1, 2);
    s->vit_key.gk = gkey;
    s->gaps = gap_cache().find(gkey);
    if (!s->gaps) {
      std::shared_ptr<GapTables> G = std::make_shared<GapTables>();
      G->rhoG_h.assign(N, 1.0);
      G->rhoZ_h.assign(N, 1.0);
      G->pair_id.assign(N, 0);
      {
      size_t cap = 1;
      int lg = 0;
      while (cap < (size_t)std::max<int64_t>(1024, N / 4)) { cap <<= 1; ++lg; }
      // multiplicative hash on the HIGH product bits: gap lengths are small integers stored as
      // doubles, so their low mantissa bits (and the low product bits) are all zero
      auto slot_of = [](uint64_t key, int bits) { return (size_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - bits)); };
      std::vector<uint64_t> keys(cap, ~0ull);
      std::vector<int> vals(cap, -1);
      std::unordered_map<std::pair<uint64_t, uint64_t>, int, th::PairHash> pairs;
      size_t used = 0;
      auto grow = [&]() {
        std::vector<uint64_t> k2(cap * 2, ~0ull);
        std::vector<int> v2(cap * 2, -1);
        for (size_t q = 0; q < cap; ++q)
          if (keys[q] != ~0ull) {
            size_t h = slot_of(keys[q], lg + 1);
            while (k2[h] != ~0ull) h = (h + 1) & (cap * 2 - 1);
            k2[h] = keys[q]; v2[h] = vals[q];
          }
        keys.swap(k2); vals.swap(v2); cap *= 2; ++lg;
      };
      for (int c = 0; c < in.nChr; ++c)
        for (int64_t t = in.chr_start[c] + 1; t < in.chr_start[c + 1]; ++t) {
          const double gap = in.posn[t] - in.posn[t - 1];      // rho_keep adds the +1.0 itself
          const uint64_t key = th::bits_of(gap);
          size_t h = slot_of(key, lg);
          while (keys[h] != ~0ull && keys[h] != key) h = (h + 1) & (cap - 1);
          if (keys[h] == ~0ull) {
            const double rG = th::rho_keep(in.posn[t - 1], in.posn[t], in.txnExpLen);
            const double rZ = th::rho_keep(in.posn[t - 1], in.posn[t], in.zStrength);
            auto pk = std::make_pair(th::bits_of(rG), th::bits_of(rZ));
            auto it = pairs.find(pk);
            if (it == pairs.end()) {
              it = pairs.emplace(pk, (int)G->pairG.size()).first;
              G->pairG.push_back(rG);
              G->pairZ.push_back(rZ);
            }
            keys[h] = key; vals[h] = it->second;
            if (++used * 2 > cap) {
              grow();
              h = slot_of(key, lg);
              while (keys[h] != key) h = (h + 1) & (cap - 1);
            }
          }
          const int id = vals[h];
          G->pair_id[t] = id;
          G->rhoG_h[t] = G->pairG[id];
          G->rhoZ_h[t] = G->pairZ[id];
        }
    }
      s->gaps = G;
      gap_cache().put(gkey, s->gaps);
    }
    trace.mark("distinct gaps (host exp)");
    TxnKey tkey;
    tkey.gk = gkey; tkey.K = K; tkey.U = s->U; tkey.Z = s->Z; tkey.h = s->h; tkey.device = s->device;
    s->txn = txn_cache().find(tkey);
    if (!s->txn) {
      DevBuf<double> dG, dZ;
      dG.alloc(N); dZ.alloc(N);
      trace.mark("  rho alloc");
      CK(cudaMemcpy(dG.p, s->gaps->rhoG_h.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dZ.p, s->gaps->rhoZ_h.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
      trace.mark("  rho memcpy");
      std::shared_ptr<TxnTable> tt = std::make_shared<TxnTable>();
      tt->device = s->device;
      tt->rec.alloc(N + 2 * kFbPadRows);
      CK(cudaMemsetAsync(tt->rec.p, 0, sizeof(TxnRec) * (N + 2 * kFbPadRows), s->stream));
      trace.mark("  txn alloc");
      txn_coeff_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s->stream>>>(dG.p, dZ.p, (int)N, K, s->U, s->Z, s->h,
                                                                           tt->rec.p + kFbPadRows);
      CK(cudaGetLastError());
      trace.mark("  txn launch");
      CK(cudaStreamSynchronize(s->stream));
      trace.mark("  txn sync");
      s->tm.kernel_launches++;
      s->txn = tt;
      txn_cache().put(tkey, s->txn);
    }
    trace.mark("  rho free");
    if (s->legacy_py) {
      s->py.alloc((size_t)N * K);
      CK(cudaMemcpy(s->py.p, in.py, sizeof(double) * N * K, cudaMemcpyHostToDevice));
      s->snp.alloc(1 + 2 * kFbPadRows);
    } else {
      if (!in.ref || !in.depth || !in.logR) return fail(TITAN_ERR_ARG, "ref, tumDepth and logR columns are required");
      rec_builder.join();
      if (bad_snp >= 0)
        return fail(TITAN_ERR_ARG, "NA/NaN in the data columns at SNP %lld (filterData removes these, R/utils.R:273-310)",
                    (long long)bad_snp);
      s->snp.alloc(N + 2 * kFbPadRows);
      CK(cudaMemset(s->snp.p, 0, sizeof(SnpRec) * (N + 2 * kFbPadRows)));
      CK(cudaMemcpy(s->snp_p(), rec.data(), sizeof(SnpRec) * N, cudaMemcpyHostToDevice));
    }
    trace.mark("SnpRec (host lgamma) + upload");
    s->alpha.alloc((size_t)(N + 2 * kFbPadRows) * s->Kp);
    {
      s->Bm.alloc((size_t)(N + 2 * kFbPadRows) * s->Kp);
      // only the slack rows need defined contents (they are copied with the tiles, never used)
      const size_t padn = (size_t)kFbPadRows * s->Kp;
      CK(cudaMemset(s->alpha.p, 0, sizeof(double) * padn));
      CK(cudaMemset(s->alpha.p + padn + (size_t)N * s->Kp, 0, sizeof(double) * padn));
      CK(cudaMemset(s->Bm.p, 0, sizeof(double) * padn));
      CK(cudaMemset(s->Bm.p + padn + (size_t)N * s->Kp, 0, sizeof(double) * padn));
    }
    trace.mark("alpha alloc");
    s->het_d.alloc(s->U);
    CK(cudaMemcpy(s->het_d.p, s->het.data(), s->U, cudaMemcpyHostToDevice));
    s->model.alloc(5 * (size_t)K + 7 * (size_t)s->U);
    s->model_h.alloc(5 * (size_t)K + 7 * (size_t)s->U);
    s->out.alloc(6 * (size_t)K + s->nChr);
    s->rhoG0.alloc((size_t)s->nChr * s->U);
    s->rhoZ0.alloc((size_t)s->nChr * s->Z);
    s->out_h.alloc(6 * (size_t)K + s->nChr + (size_t)s->nChr * (s->U + s->Z));
    s->rerun.alloc(2 * kRoundSlots);
    s->rerun_h.alloc(2 * kRoundSlots);
    build_chunks(s);
    trace.mark("small buffers + chunks");
    if (!s->legacy_py && K <= 256)
      s->vit_builder = std::thread([s]() {
        // the whole preparation of the decode runs beside the EM: ~40 ms of host libm for the tables, then their upload
        // (32 K bytes per distinct transition pair: 50-80 MB) and the decode's device buffers; titan_viterbi only joins
        try {
          CK(cudaSetDevice(s->device));
          s->vt = vit_cache().find(s->vit_key);       // another solution of the sweep with the same K built them already
          if (!s->vt) build_viterbi_tables(s);
          viterbi_device_setup(s);
        } catch (const CudaFail &f) {
          s->vit_err = f.code;
          s->vit_errmsg = g_err;
        } catch (const std::bad_alloc &) {
          s->vit_err = TITAN_ERR_ARG;
          s->vit_errmsg = "host allocation failed while preparing the Viterbi tables";
        }
      });
  } catch (const CudaFail &f) {
    return f.code;
  } catch (const std::bad_alloc &) {
    return fail(TITAN_ERR_ARG, "host allocation failed");
  }
  *out = S.release();
  return TITAN_OK;
}

extern "C" int titan_solver_create(const titan_data_desc *d, titan_solver **out) {
  if (!d) return fail(TITAN_ERR_ARG, "titan_solver_create: desc is NULL");
  SolverInit in{d->N, d->nChr, d->chr_start, d->posn, d->ref, d->tumDepth, d->logR, nullptr,
                d->txnExpLen, d->txnZstrength * d->txnExpLen, d->U, d->Z, d->ZS, d->U};
  in.alleleModel = d->alleleModel;
  return solver_create_impl(in, out);
}

extern "C" void titan_solver_destroy(titan_solver *s) {
  PhaseTrace tr_all;
  if (!s) return;
  cudaSetDevice(s->device);
  delete s;
  tr_all.mark("destroy: total (device buffers last)");
}

extern "C" int titan_solver_shape(const titan_solver *s, int64_t *N, int32_t *nChr, int32_t *U, int32_t *Z) {
  if (!s) return fail(TITAN_ERR_ARG, "titan_solver_shape: NULL solver");
  if (N) *N = s->N;
  if (nChr) *nChr = s->nChr;
  if (U) *U = s->U;
  if (Z) *Z = s->Z;
  return TITAN_OK;
}

extern "C" int titan_solver_configure(titan_solver *s, int chunk_len, int warmup_len, double rel_tol, int max_rounds) {
  if (!s) return fail(TITAN_ERR_ARG, "solver is NULL");
  try {
    s->chunk_len = chunk_len;
    s->warm = std::max(1, warmup_len);
    if (rel_tol > 0) s->tol = rel_tol;
    s->max_rounds = std::max(0, max_rounds);
    CK(cudaSetDevice(s->device));
    build_chunks(s);
    s->state_valid = false;
  } catch (const CudaFail &f) {
    return f.code;
  }
  return TITAN_OK;
}

extern "C" int titan_solver_set_stream(titan_solver *s, void *cuda_stream) {
  if (!s) return fail(TITAN_ERR_ARG, "solver is NULL");
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);      // nothing of this solver may still be in flight on the old stream
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  if (cuda_stream) {
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(TITAN_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    s->own_stream = true;
  }
  return TITAN_OK;
}

extern "C" int titan_release_cached_memory(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(TITAN_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  vit_cache().clear();                                // shared tables no live solver holds go back to the pool first
  txn_cache().clear();
  cudaMemPool_t pool = titan_pool(dev);
  if (pool) {
    cudaDeviceSynchronize();
    e = cudaMemPoolTrimTo(pool, 0);
  }
  if (e != cudaSuccess) return fail(TITAN_ERR_CUDA, "cudaMemPoolTrimTo: %s", cudaGetErrorString(e));
  return TITAN_OK;
}

extern "C" int titan_solver_timing(const titan_solver *s, titan_timing *t) {
  if (!s || !t) return fail(TITAN_ERR_ARG, "NULL argument");
  *t = s->tm;
  return TITAN_OK;
}

// ---------------------------------------------------------------------------------------------
// chromosome shards: whole-genome layout of the per-E-step outputs, all-reduced over the ranks of one solution
// ---------------------------------------------------------------------------------------------
__global__ void shard_gather_kernel(const double *__restrict__ out, const double *__restrict__ rhoG0,
                                    const double *__restrict__ rhoZ0, const int *__restrict__ chr_global, int nChr, int nT,
                                    int width, int U, int Z, double *__restrict__ g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < width) { g[i] = out[i]; return; }
  int j = i - width;
  if (j < nChr) { g[width + chr_global[j]] = out[width + j]; return; }
  j -= nChr;
  if (j < nChr * U) { g[width + nT + chr_global[j / U] * U + j % U] = rhoG0[j]; return; }
  j -= nChr * U;
  if (j < nChr * Z) g[width + nT + nT * U + chr_global[j / Z] * Z + j % Z] = rhoZ0[j];
}

extern "C" int titan_solver_set_comm(titan_solver *s, titan_comm *comm, int nChr_total, const int32_t *chr_global) {
  if (!s) return fail(TITAN_ERR_ARG, "titan_solver_set_comm: NULL solver");
  if (s->legacy_py) return fail(TITAN_ERR_ARG, "titan_solver_set_comm: solver was created for materialised emissions");
  try {
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->state_valid = false;
    if (!comm) {
      s->comm = nullptr;
      s->out_h.alloc(6 * (size_t)s->K + s->nChr + (size_t)s->nChr * (s->U + s->Z));
      return TITAN_OK;
    }
    if (!chr_global || nChr_total < s->nChr) return fail(TITAN_ERR_ARG, "titan_solver_set_comm: chr_global[nChr] and nChr_total >= nChr are required");
    if (comm->device != s->device) return fail(TITAN_ERR_ARG, "titan_solver_set_comm: communicator and solver live on different devices");
    for (int c = 0; c < s->nChr; ++c)
      if (chr_global[c] < 0 || chr_global[c] >= nChr_total || (c > 0 && chr_global[c] <= chr_global[c - 1]))
        return fail(TITAN_ERR_ARG, "titan_solver_set_comm: chr_global must be ascending indices below nChr_total");
    s->chr_global.assign(chr_global, chr_global + s->nChr);
    s->nChrTotal = nChr_total;
    s->chr_global_d.alloc(std::max(1, s->nChr));
    CK(cudaMemcpy(s->chr_global_d.p, s->chr_global.data(), sizeof(int) * s->nChr, cudaMemcpyHostToDevice));
    const size_t len = 6 * (size_t)s->K + (size_t)nChr_total * (1 + s->U + s->Z);
    s->gvec.alloc(len);
    s->out_h.alloc(len);
    s->comm = comm;
  } catch (const CudaFail &f) {
    return f.code;
  }
  return TITAN_OK;
}

// ---------------------------------------------------------------------------------------------
// model constants
// ---------------------------------------------------------------------------------------------
static ModelConsts model_layout(const titan_solver *s, double *base) {
  const size_t K = s->K, U = s->U;
  ModelConsts M;
  M.lmu = base;
  M.l1mu = base + K;
  M.muC = base + 2 * K;
  M.pi = base + 3 * K;
  M.logpi = base + 4 * K;
  M.cN = base + 5 * K;
  M.i2v = base + 5 * K + U;
  M.tv = base + 5 * K + 2 * U;
  M.gR = base + 5 * K + 3 * U;
  M.gX = base + 5 * K + 4 * U;
  M.vR = base + 5 * K + 5 * U;
  M.sq = base + 5 * K + 6 * U;
  M.l0 = s->model_l0; M.rho2 = s->model_rho2;
  M.gauss = s->gauss ? 1 : 0;
  return M;
}

// host part of the emission (R/hmmClonal.R:579-603,617-634): K logs of the mixture means and the
// U Gaussian constants, evaluated with the host libm so the Viterbi emission is bit-exact.
static int upload_model(titan_solver *s, const titan_params *p, double *muR_out, double *muC_out) {
  const int U = s->U, Z = s->Z, K = s->K;
  double *hb = s->model_h.p;
  ModelConsts H = model_layout(s, hb);
  double *lmu = (double *)H.lmu, *l1mu = (double *)H.l1mu, *muC = (double *)H.muC, *pi = (double *)H.pi,
         *logpi = (double *)H.logpi, *cN = (double *)H.cN, *i2v = (double *)H.i2v, *tv = (double *)H.tv;
  if (p->piG && p->piZ) {
    for (int z = 0; z < Z; ++z)
      for (int u = 0; u < U; ++u) {
        pi[u + z * U] = p->piG[u] * p->piZ[z];            // R/hmmClonal.R:196
        logpi[u + z * U] = std::log(pi[u + z * U]);       // :210 log(piGiZi[c, ])
      }
  }
  if (!s->legacy_py) {
    if (!p->s || !p->var || !p->rt || !p->ct) return fail(TITAN_ERR_ARG, "titan_params: s, var, rt and ct are required");
    std::vector<double> muR((size_t)K);
    th::mixture_means(U, Z, p->rt, p->rn, p->n, p->s, p->ct, p->phi, muR.data(), muC);
    if (s->gauss) {
      // bivariateNormalpdflog (R/hmmClonal.R:550-559) with var1 = varR (allelic), var2 = var (logR): the constants in
      // R's association order for the exact (Viterbi) evaluation, and folded coefficients for the E-step
      if (!p->varR) return fail(TITAN_ERR_ARG, "titan_params: varR is required by the Gaussian allele model");
      const double rho = p->corRho;
      if (!(rho > -1 && rho < 1)) return fail(TITAN_ERR_NUMERIC, "corRho=%g outside (-1, 1)", rho);
      double *gR = (double *)H.gR, *gX = (double *)H.gX, *vR = (double *)H.vR, *sq = (double *)H.sq;
      const double l0 = 1 / (2 * (1 - rho * rho));
      s->model_l0 = l0;
      s->model_rho2 = 2 * rho;
      for (int j = 0; j < K; ++j) {
        if (!std::isfinite(muR[j]) || !std::isfinite(muC[j]))
          return fail(TITAN_ERR_NUMERIC, "mixture mean out of range at state %d (muR=%g muC=%g)", j, muR[j], muC[j]);
        lmu[j] = muR[j];
        l1mu[j] = 0.0;
      }
      for (int u = 0; u < U; ++u) {
        const double var1 = p->varR[u], var2 = p->var[u];
        if (!(var1 > 0) || !(var2 > 0))
          return fail(TITAN_ERR_NUMERIC, "non-positive variance at state %d (varR=%g var=%g)", u, var1, var2);
        cN[u] = -std::log((2 * th::kPi) * std::sqrt((var1 * var2) * (1 - rho * rho)));
        tv[u] = var2;
        vR[u] = var1;
        sq[u] = std::sqrt(var1 * var2);
        i2v[u] = l0 / var2;
        gR[u] = l0 / var1;
        gX[u] = l0 * (2 * rho) / sq[u];
      }
    } else {
    for (int j = 0; j < K; ++j) {
      if (!(muR[j] > 0) || !(muR[j] <= 1) || !std::isfinite(muC[j]))
        return fail(TITAN_ERR_NUMERIC, "mixture mean out of range at state %d (muR=%g muC=%g)", j, muR[j], muC[j]);
      lmu[j] = std::log(muR[j]);
      l1mu[j] = std::log(1 - muR[j]);
    }
    for (int u = 0; u < U; ++u) {
      if (!(p->var[u] > 0)) return fail(TITAN_ERR_NUMERIC, "non-positive variance var[%d]=%g", u, p->var[u]);
      cN[u] = std::log(1 / (std::sqrt(p->var[u]) * std::sqrt(2 * th::kPi)));
      tv[u] = 2 * p->var[u];
      i2v[u] = 1.0 / tv[u];
    }
    }
    if (muR_out) std::copy(muR.begin(), muR.end(), muR_out);
    if (muC_out) std::copy(muC, muC + K, muC_out);
  }
  CK(cudaMemcpyAsync(s->model.p, hb, sizeof(double) * (5 * (size_t)K + (s->gauss ? 7 : 3) * (size_t)U), cudaMemcpyHostToDevice,
                     s->stream));
  return TITAN_OK;
}

// ---------------------------------------------------------------------------------------------
// E-step launch sequence
// ---------------------------------------------------------------------------------------------
// fb2 engine: emission pass, round 0, verification rounds (both directions in one launch per round, batches of
// rounds per host synchronisation), final posterior pass.  Events: ev[0] start, ev[6] after the emission pass,
// ev[7] after round 0, ev[1] after the last round, ev[9] after the final pass.
static int estep_fb2(titan_solver *s, bool post, bool want_loggamma, bool posterior_only, int *rounds_out, int64_t *runs_out) {
  const int U = s->U, K = s->K;
  Fb2Args A{};
  A.Bm = s->Bm_p();
  A.txn = s->txn_p();
  A.snp = s->snp_p();
  A.chunks = s->chunks.p;
  A.nChunks = s->nChunks;
  A.U = U; A.K = K; A.Kp = s->Kp;
  A.warm = s->warm;
  A.tol = s->tol;
  A.het = s->het_d.p;
  A.pi = model_layout(s, s->model.p).pi;
  A.alpha = s->alpha_p();
  A.startF = s->startf.p; A.endF = s->endf.p; A.extF = s->extf.p;
  A.startB = s->startb.p; A.endB = s->endb.p; A.extB = s->extb.p;
  A.chunk_ll = s->chunk_ll.p;
  A.subB = s->subb.p;
  A.rerun = s->rerun.p;
  {
    const size_t nc = s->nChunks;
    A.dirtyF = s->pflags.p; A.dirtyB = s->pflags.p + nc; A.solvedF = s->pflags.p + 2 * nc; A.solvedB = s->pflags.p + 3 * nc;
    A.probedF = s->pflags.p + 4 * nc; A.probedB = s->pflags.p + 5 * nc;
    A.probe_window = 0;
    A.probeJF = s->pints.p; A.probeJB = s->pints.p + (kProbes - 1) * nc;
    A.probeEF = s->pints.p + 2 * (kProbes - 1) * nc; A.probeEB = s->pints.p + (3 * kProbes - 2) * nc;
    A.probeF = s->pvec.p; A.probeB = s->pvec.p + kProbes * nc * K;
    A.solF = s->pvec.p + 2 * kProbes * nc * K; A.solB = s->pvec.p + (2 * kProbes + 1) * nc * K;
    A.chr_chunk0 = s->chr_chunk0.p;
    A.nChr = s->nChr;
    A.use_solved = 0;
  }
  A.part = s->part.p;
  A.final_mode = 0;
  A.ver = s->ver.p; A.fin_ver = s->ver.p + s->nChunks;
  A.rhoG0 = s->rhoG0.p; A.rhoZ0 = s->rhoZ0.p;
  A.rhoG = (post && !s->legacy_py) ? s->rhoG.p : nullptr;
  A.rhoZ = (post && !s->legacy_py) ? s->rhoZ.p : nullptr;
  A.loggamma = (s->legacy_py && want_loggamma) ? s->loggamma.p : nullptr;
  cudaStream_t st = s->stream;
  int rounds = 1;
  int64_t runs = 0;
  bool spec = false, spec_launched_any = false;
  if (!posterior_only) {
    EmArgs E{};
    E.snp = s->snp_p();
    E.py = s->legacy_py ? s->py.p : nullptr;
    E.M = model_layout(s, s->model.p);
    E.U = U; E.K = K; E.Kp = s->Kp; E.N = s->N;
    E.Bm = s->Bm_p();
    CK(fb2_launch_emission(E, s->Z, st));
    CK(cudaEventRecord(s->ev[6], st));
    const int hard_cap = std::min(kRoundSlots - 8, s->max_rounds > 0 ? s->max_rounds : 2 * s->maxChunksPerChr + 4);
    CK(cudaMemsetAsync(s->rerun.p, 0, sizeof(unsigned int) * (size_t)(hard_cap + 8), st));
    A.round = 0;
    CK(fb2_launch_round(A, s->Z, st));
    CK(cudaEventRecord(s->ev[7], st));
    s->tm.kernel_launches += 2;
    runs = 2 * (int64_t)s->nChunks;
    bool done = s->maxChunksPerChr <= 1;            // plain sequential recursion: nothing to verify
    int round = 0;
    unsigned int *ctr_h = s->rerun_h.p;
    const bool probes = s->probes != 0 && K >= 3;
    // Speculative final pass: the cycles after round 0 are latency-bound (one recursion warp per dirty chunk, most SMs
    // idle), the final pass is throughput work.  Once the first cycle has decided which chunks it re-runs, every other
    // chunk gets its final pass on a second stream; a chunk that runs again later bumps its counter and is redone by
    // the closing pass.  Results are identical: a chunk's final pass depends only on that chunk's own verified inputs.
    spec = probes && !done && s->spec_final != 0 && !s->legacy_py;
    if (spec) {
      if (!s->stream2) CK(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
      for (auto &e : s->ev_spec)
        if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CK(cudaMemsetAsync(s->ver.p, 0, sizeof(unsigned int) * (size_t)s->nChunks, st));
      CK(cudaMemsetAsync(s->ver.p + s->nChunks, 0xff, sizeof(unsigned int) * (size_t)s->nChunks, st));
    }
    bool &spec_launched = spec_launched_any;
    while (!done) {
      // probe mode: a cycle is boundary test + probe runs, chain solve, one re-run of the dirty chunks (3 launches)
      const int queued = std::max(1, std::min(probes ? std::max(1, s->round_batch / 2) : s->round_batch, hard_cap - round));
      for (int q = 1; q <= queued; ++q) {
        A.round = round + q;
        A.use_solved = 0;
        if (probes && (s->plain_after <= 0 || round + q <= s->plain_after)) {
          A.use_solved = 1;
          A.probe_window = (round + q >= 2) ? s->probe_window : s->probe_window1;   // cycle 1 tests hundreds of chunks
          CK(fb2_launch_probe(A, s->Z, st));
          CK(fb2_launch_solve(A, s->Z, st));
          s->tm.kernel_launches += 2;
          if (spec) {     // this cycle has its re-run set: every other chunk without a valid final pass goes to the second stream
            CK(cudaEventRecord(s->ev_spec[0], st));
            CK(cudaStreamWaitEvent(s->stream2, s->ev_spec[0], 0));
            Fb2Args F = A;
            F.final_mode = 1;
            CK(fb2_launch_final(F, s->Z, post, false, s->stream2));
            CK(cudaEventRecord(s->ev_spec[1], s->stream2));
            s->tm.kernel_launches++;
            spec_launched = true;
          }
        }
        CK(fb2_launch_round(A, s->Z, st));
        s->tm.kernel_launches++;
      }
      CK(cudaMemcpyAsync(ctr_h + round + 1, s->rerun.p + round + 1, sizeof(unsigned int) * (size_t)queued, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      for (int q = 1; q <= queued; ++q) {
        ++round;
        if (ctr_h[round] == 0) { done = true; break; }
        runs += ctr_h[round];
      }
      if (!done && round >= hard_cap) {
        if (s->stream2) cudaStreamSynchronize(s->stream2);      // no speculative pass may outlive the failed call
        return fail(TITAN_ERR_NUMERIC, "chunk boundaries did not reach a fixed point in %d rounds", hard_cap);
      }
    }
    rounds = round + 1;
    if (getenv("TITAN_TRACE") && atoi(getenv("TITAN_TRACE")) > 1) {
      fprintf(stderr, "[titan trace] chunk runs per round (both directions, %d chunks):", s->nChunks);
      for (int r = 1; r <= round; ++r) fprintf(stderr, " %u", ctr_h[r]);
      fprintf(stderr, "\n");
    }
  } else {
    CK(cudaEventRecord(s->ev[6], st));
    CK(cudaEventRecord(s->ev[7], st));
  }
  CK(cudaEventRecord(s->ev[1], st));
  if (spec_launched_any) {
    CK(cudaStreamWaitEvent(st, s->ev_spec[1], 0));
    A.final_mode = 2;                               // closing pass: chunks that ran again since their speculative pass
  }
  CK(cudaEventRecord(s->ev[8], st));
  CK(fb2_launch_final(A, s->Z, post && !s->legacy_py, s->legacy_py, st));
  CK(cudaEventRecord(s->ev[9], st));
  s->tm.kernel_launches++;
  runs += s->nChunks;
  *rounds_out = rounds;
  *runs_out = runs;
  return TITAN_OK;
}

// posterior_only: the parameters are those of the previous call, whose forward variables and verified
// backward boundaries are still resident -- only the final backward pass is run again, now writing
// the N-sized marginals (what runEMclonalCN returns as rhoG / rhoZ of its last E-step).
static int estep_impl(titan_solver *s, const titan_params *p, titan_estep_out *o, bool want_post, double *loggamma_out,
                      bool posterior_only = false) {
  const int U = s->U, Z = s->Z, K = s->K;
  int nChr = s->nChr;
  const int64_t N = s->N;
  CK(cudaSetDevice(s->device));
  int rc = TITAN_OK;
  posterior_only = posterior_only && s->state_valid;
  s->state_valid = false;
  if (!posterior_only) {
    rc = upload_model(s, p, o ? o->muR : nullptr, o ? o->muC : nullptr);
    if (rc != TITAN_OK) return rc;
  }
  const bool post = want_post || (o && (o->rhoG || o->rhoZ));
  if (post && !s->legacy_py) {
    if (s->rhoG.n < (size_t)N * U) s->rhoG.alloc((size_t)N * U);
    if (s->rhoZ.n < (size_t)N * Z) s->rhoZ.alloc((size_t)N * Z);
  }
  if (s->legacy_py && loggamma_out && s->loggamma.n < (size_t)N * K) s->loggamma.alloc((size_t)N * K);

  CK(cudaEventRecord(s->ev[0], s->stream));
  int fb2_rounds = 0;
  int64_t fb2_runs = 0;
  rc = estep_fb2(s, post, loggamma_out != nullptr, posterior_only, &fb2_rounds, &fb2_runs);
  if (rc != TITAN_OK) return rc;
  CK(cudaEventRecord(s->ev[2], s->stream));
  const int fr = fb2_rounds, br = fb2_rounds;
  const int64_t fruns = fb2_runs, bruns = 0;
  const int width = 6 * K;
  {
    reduce_kernel<<<(width + 31) / 32 + 1, 256, 0, s->stream>>>(s->part.p, s->nChunks, width, s->chunk_ll.p, s->chr_chunk0.p, nChr,
                                                                  s->out.p);
    s->tm.kernel_launches += 1;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(s->ev[3], s->stream));
  double *oh = s->out_h.p;
  if (s->comm) {
    // whole-genome layout, zero where another rank holds the chromosome, summed over the ranks on the device
    const int nT = s->nChrTotal;
    const size_t len = (size_t)width + (size_t)nT * (1 + U + Z);
    const int items = width + nChr * (1 + U + Z);
    CK(cudaMemsetAsync(s->gvec.p, 0, sizeof(double) * len, s->stream));
    shard_gather_kernel<<<(items + 255) / 256, 256, 0, s->stream>>>(s->out.p, s->rhoG0.p, s->rhoZ0.p, s->chr_global_d.p, nChr, nT,
                                                                    width, U, Z, s->gvec.p);
    CK(cudaGetLastError());
    s->tm.kernel_launches += 1;
    std::string err;
    if (!comm_allreduce_sum_f64(s->comm, s->gvec.p, len, s->stream, err)) return fail(TITAN_ERR_CUDA, "%s", err.c_str());
    CK(cudaEventRecord(s->ev[3], s->stream));      // the collective is part of a sharded E-step's device time
    CK(cudaMemcpyAsync(oh, s->gvec.p, sizeof(double) * len, cudaMemcpyDeviceToHost, s->stream));
    nChr = nT;                      // the outputs below are whole-genome sized
  } else {
    CK(cudaMemcpyAsync(oh, s->out.p, sizeof(double) * (width + nChr), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(oh + width + nChr, s->rhoG0.p, sizeof(double) * (size_t)nChr * U, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(oh + width + nChr + (size_t)nChr * U, s->rhoZ0.p, sizeof(double) * (size_t)nChr * Z,
                       cudaMemcpyDeviceToHost, s->stream));
  }
  CK(cudaStreamSynchronize(s->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1])); s->tm.last_fwd_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[1], s->ev[2])); s->tm.last_bwd_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3])); s->tm.last_reduce_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[0], s->ev[3])); s->tm.last_estep_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[6], s->ev[7])); s->tm.last_fwd_r0_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[8], s->ev[9])); s->tm.last_bwd_r0_ms = ms;
  s->tm.engine = s->engine;
  s->tm.last_emission_ms = 0;
  s->tm.last_rounds_ms = 0;
  CK(cudaEventElapsedTime(&ms, s->ev[0], s->ev[6])); s->tm.last_emission_ms = ms;
  CK(cudaEventElapsedTime(&ms, s->ev[7], s->ev[1])); s->tm.last_rounds_ms = ms;
  s->tm.last_fwd_rounds = fr; s->tm.last_bwd_rounds = br;
  s->tm.last_fwd_chunk_runs = fruns; s->tm.last_bwd_chunk_runs = bruns;
  if (o) {
    if (o->stats) std::copy(oh, oh + width, o->stats);
    if (o->loglik_chr) std::copy(oh + width, oh + width + nChr, o->loglik_chr);
    if (o->rhoG0) std::copy(oh + width + nChr, oh + width + nChr + (size_t)nChr * U, o->rhoG0);
    if (o->rhoZ0)
      std::copy(oh + width + nChr + (size_t)nChr * U, oh + width + nChr + (size_t)nChr * (U + Z), o->rhoZ0);
    if (o->rhoG && !s->legacy_py) CK(cudaMemcpy(o->rhoG, s->rhoG.p, sizeof(double) * (size_t)N * U, cudaMemcpyDeviceToHost));
    if (o->rhoZ && !s->legacy_py) CK(cudaMemcpy(o->rhoZ, s->rhoZ.p, sizeof(double) * (size_t)N * Z, cudaMemcpyDeviceToHost));
  }
  if (loggamma_out && s->legacy_py)
    CK(cudaMemcpy(loggamma_out, s->loggamma.p, sizeof(double) * (size_t)N * K, cudaMemcpyDeviceToHost));
  s->state_valid = !s->legacy_py;
  return TITAN_OK;
}

extern "C" int titan_estep(titan_solver *s, const titan_params *p, titan_estep_out *out) {
  if (!s || !p) return fail(TITAN_ERR_ARG, "titan_estep: NULL argument");
  if (s->legacy_py) return fail(TITAN_ERR_ARG, "titan_estep: solver was created for materialised emissions");
  if (!p->piG || !p->piZ) return fail(TITAN_ERR_ARG, "titan_estep: piG and piZ are required");
  try {
    return estep_impl(s, p, out, false, nullptr);
  } catch (const CudaFail &f) {
    return f.code;
  }
}

// ---------------------------------------------------------------------------------------------
// Viterbi
// ---------------------------------------------------------------------------------------------
// ---- tiled back-trace ----
static std::vector<int> bt_tile_plan(const int *cs, int nChr, int K, int &rows) {
  rows = K <= 128 ? 256 : 128;
  std::vector<int> tf(nChr + 1, 0);
  for (int c = 0; c < nChr; ++c) {
    const int n = std::max(cs[c + 1] - cs[c] - 1, 0);      // psi rows cs+1 .. ce-1
    tf[c + 1] = tf[c] + (n + rows - 1) / rows;
  }
  return tf;
}

static void launch_backtrace(const VitArgs &A, int nTiles, cudaStream_t st) {
  const size_t smem = (size_t)A.bt_rows * A.K + 64 + A.bt_rows;
  if (nTiles > 0) viterbi_bt_compose_kernel<<<nTiles, 256, smem, st>>>(A);
  viterbi_bt_link_kernel<<<(A.nChr + 31) / 32, 32, 0, st>>>(A);
  if (nTiles > 0) viterbi_bt_walk_kernel<<<nTiles, 256, smem, st>>>(A);
}

// exact log-transition tables, one per distinct (rhoG, rhoZ) pair (host libm, sequential row sums);
// independent per pair, so they are spread over the host threads.  Parameter-independent: started in
// the background at solver creation and joined by the first decode.
static void build_viterbi_tables(titan_solver *s) {
  const int K = s->K;
  const size_t np = std::max<size_t>(1, s->gaps->pairG.size());
  s->vit_tab.assign(np * 4 * (size_t)K, 0.0);
  const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), (unsigned)((s->gaps->pairG.size() + 255) / 256)));
  std::vector<unsigned char> cls;                       // class of every (source, destination), once for all pairs
  th::structured_class_matrix(K, s->U, s->het.data(), cls);
  const unsigned char *cl = cls.data();
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < nt; ++w)
    pool.emplace_back([s, w, nt, K, cl]() {
      for (size_t q = w; q < s->gaps->pairG.size(); q += nt)
        th::viterbi_log_table_rows8(K, cl, s->gaps->pairG[q], s->gaps->pairZ[q], s->vit_tab.data() + q * 4 * (size_t)K);
    });
  for (auto &t : pool) t.join();
}

static int viterbi_prepare(titan_solver *s) {
  if (s->vit_ready) return TITAN_OK;
  if (s->K > 256) return fail(TITAN_ERR_UNSUPPORTED, "Viterbi back-pointers are 8-bit: K=%d > 256", s->K);
  if (s->vit_builder.joinable()) s->vit_builder.join();
  if (s->vit_err != TITAN_OK) return fail(s->vit_err, "%s", s->vit_errmsg.c_str());
  if (s->vit_ready) return TITAN_OK;                 // the builder thread finished the job
  if (!s->vt && s->vit_tab.empty()) build_viterbi_tables(s);   // materialised emissions (legacy entry): no builder thread
  viterbi_device_setup(s);
  return TITAN_OK;
}

// device side of the preparation: chromosome offsets, table ids, log-transition tables, back-pointer and back-trace
// buffers.  Runs on the builder thread (its own cudaStreamPerThread) or, for the legacy entry, on the caller's.
static void viterbi_device_setup(titan_solver *s) {
  const int64_t N = s->N;
  const int K = s->K;
  std::vector<int> cs32(s->nChr + 1);
  for (int c = 0; c <= s->nChr; ++c) cs32[c] = (int)s->chr_start[c];
  s->chr_start32.alloc(s->nChr + 1);
  CK(cudaMemcpy(s->chr_start32.p, cs32.data(), sizeof(int) * (s->nChr + 1), cudaMemcpyHostToDevice));
  if (!s->vt) {
    const std::vector<double> &tab = s->vit_tab;
    const std::vector<int> &id = s->gaps->pair_id;
    std::shared_ptr<VitTables> vt = std::make_shared<VitTables>();
    vt->device = s->device;
    // the warp-reduction kernel relies on: same-genotype classes never below the other-genotype classes of the same row
    vt->monotone = true;
    for (size_t r = 0; r + 3 < tab.size() && vt->monotone; r += 4) vt->monotone = tab[r + 2] >= tab[r] && tab[r + 3] >= tab[r + 1];
    vt->txn_id.alloc(N);
    CK(cudaMemcpy(vt->txn_id.p, id.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    vt->ltab.alloc(tab.size());
    CK(cudaMemcpy(vt->ltab.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    std::vector<double>().swap(s->vit_tab);
    s->vt = vt;
    if (!s->legacy_py) vit_cache().put(s->vit_key, s->vt);
  }
  s->vit_monotone = s->vt->monotone;
  s->psi.alloc((size_t)N * K + 64);       // slack: the back-trace stages tiles with 16-byte loads
  {
    const std::vector<int> tf = bt_tile_plan(cs32.data(), s->nChr, K, s->bt_rows);
    s->nTiles = tf.back();
    s->tile_first.alloc(tf.size());
    CK(cudaMemcpy(s->tile_first.p, tf.data(), sizeof(int) * tf.size(), cudaMemcpyHostToDevice));
    s->bt_comp.alloc((size_t)std::max(1, s->nTiles) * K);
    s->bt_entry.alloc((size_t)std::max(1, s->nTiles));
  }
  s->path.alloc(N);
  s->vit_ready = true;
}

// ---- Viterbi forward-kernel selection ----
constexpr int kVitClusterAutoMaxZ = 4;
static bool vit_fast_cfg(int K, int &P, int &PER) {
  int want = 0;
  if (const char *e = getenv("TITAN_VIT_P")) want = atoi(e);
  if (K <= 40) { P = 8; PER = 5; }
  else if (K <= 80) {
    if (want == 8) { P = 8; PER = 10; } else { P = 4; PER = 20; }
  }
  else if (K <= 104) { P = 4; PER = 32; }
  else if (K <= 128) { P = 2; PER = 64; }
  else return false;
  return true;
}

template <int PER, int P, int MAXT>
static void launch_vit_fast_t(const VitArgs &A, int nChr, int threads, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(viterbi_fast_kernel<PER, P, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  viterbi_fast_kernel<PER, P, MAXT><<<nChr, threads, smem, st>>>(A);
}

static void launch_vit_fast(const VitArgs &A, int nChr, int K, int P, int PER, cudaStream_t st) {
  const int threads = ((K * P + 31) / 32) * 32;
  const size_t smem = sizeof(double) * (2 * kVitVS + kVitMaxK) + VitRing::bytes(K);
  switch (PER) {
    case 5: launch_vit_fast_t<5, 8, 320>(A, nChr, threads, smem, st); break;
    case 10: launch_vit_fast_t<10, 8, 640>(A, nChr, threads, smem, st); break;
    case 20: launch_vit_fast_t<20, 4, 320>(A, nChr, threads, smem, st); break;
    case 32: launch_vit_fast_t<32, 4, 416>(A, nChr, threads, smem, st); break;
    default: launch_vit_fast_t<64, 2, 256>(A, nChr, threads, smem, st); break;
  }
}

template <int Z, int H, int UP>
static void launch_vit_cluster_t(const VitArgs &A, int nChr, cudaStream_t st) {
  constexpr int W = Z * H;
  const size_t smem = sizeof(double) * W * 33 * 4 + 16 * (size_t)(2 * 2 * W * 32) + sizeof(double) * (size_t)(A.K + 2) +
                      VitRing::bytes(A.K);
  if (smem > 48 * 1024) cudaFuncSetAttribute(viterbi_cluster_kernel<Z, H, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  viterbi_cluster_kernel<Z, H, UP><<<nChr, 32 * W, smem, st>>>(A);
}

template <int Z>
static void launch_vit_cluster_z(const VitArgs &A, int nChr, cudaStream_t st) {
  // slices per cluster: measured on B200 at 1 M SNPs -- four warps pay off only for a single cluster (23.5 vs 26.6 ms);
  // with Z >= 2 the longer gather and the wider barrier cost more than the shorter slice scan saves
  if (Z == 1 && A.U <= 28) { launch_vit_cluster_t<Z, (Z == 1 ? 4 : 2), (Z == 1 ? 7 : 13)>(A, nChr, st); return; }
  if (A.U <= 26) launch_vit_cluster_t<Z, 2, 13>(A, nChr, st);
  else launch_vit_cluster_t<Z, 2, 16>(A, nChr, st);
}

static void launch_vit_cluster(const VitArgs &A, int nChr, cudaStream_t st) {
  switch (A.Z) {
    case 1: launch_vit_cluster_z<1>(A, nChr, st); break;
    case 2: launch_vit_cluster_z<2>(A, nChr, st); break;
    case 3: launch_vit_cluster_z<3>(A, nChr, st); break;
    case 4: launch_vit_cluster_z<4>(A, nChr, st); break;
    case 5: launch_vit_cluster_z<5>(A, nChr, st); break;
    case 6: launch_vit_cluster_z<6>(A, nChr, st); break;
    case 7: launch_vit_cluster_z<7>(A, nChr, st); break;
    default: launch_vit_cluster_z<8>(A, nChr, st); break;
  }
}

template <int Z, int GS>
static void launch_vit_redux_zg(const VitArgs &A, int nChr, cudaStream_t st) {
  const size_t smem = 16 * (size_t)(2 * 2 * Z * 32 * GS) + sizeof(double) * (size_t)(A.K + 2) + VitRing::bytes(A.K);
  if (smem > 48 * 1024) cudaFuncSetAttribute(viterbi_redux_kernel<Z, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  viterbi_redux_kernel<Z, GS><<<nChr, 32 * (Z + 1), smem, st>>>(A);
}
template <int Z>
static void launch_vit_redux_z(const VitArgs &A, int nChr, cudaStream_t st) {
  if (A.U > 32) launch_vit_redux_zg<Z, 2>(A, nChr, st);
  else launch_vit_redux_zg<Z, 1>(A, nChr, st);
}

static void launch_vit_redux(const VitArgs &A, int nChr, cudaStream_t st) {
  switch (A.Z) {
    case 1: launch_vit_redux_z<1>(A, nChr, st); break;
    case 2: launch_vit_redux_z<2>(A, nChr, st); break;
    case 3: launch_vit_redux_z<3>(A, nChr, st); break;
    case 4: launch_vit_redux_z<4>(A, nChr, st); break;
    case 5: launch_vit_redux_z<5>(A, nChr, st); break;
    case 6: launch_vit_redux_z<6>(A, nChr, st); break;
    case 7: launch_vit_redux_z<7>(A, nChr, st); break;
    default: launch_vit_redux_z<8>(A, nChr, st); break;
  }
}

static void launch_vit_dense(const VitArgs &A, int nChr, int K, bool from_py, cudaStream_t st) {
  int P = 4;                                   // threads per destination state (power of two)
  while (P > 1 && K * P > 1024) P >>= 1;
  const int threads = std::max(((K * P + 31) / 32) * 32, ((2 * K + 31) / 32) * 32);   // >= 2K so two loads per thread cover a 4K-entry table
  const size_t smem = sizeof(double) * 13 * (size_t)K + sizeof(unsigned short) * (size_t)((K + P - 1) / P) * threads;
  if (smem > 48 * 1024) {
    if (from_py) cudaFuncSetAttribute(viterbi_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    else cudaFuncSetAttribute(viterbi_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  if (from_py) viterbi_forward_kernel<true><<<nChr, threads, smem, st>>>(A, P);
  else viterbi_forward_kernel<false><<<nChr, threads, smem, st>>>(A, P);
}

static int viterbi_impl(titan_solver *s, const titan_params *p, int32_t *path_out) {
  CK(cudaSetDevice(s->device));
  int rc = viterbi_prepare(s);
  if (rc != TITAN_OK) return rc;
  rc = upload_model(s, p, nullptr, nullptr);
  if (rc != TITAN_OK) return rc;
  VitArgs A{};
  A.snp = s->snp_p();
  A.chr_start = s->chr_start32.p;
  A.txn_id = s->vt->txn_id.p;
  A.ltab = s->vt->ltab.p;
  A.M = model_layout(s, s->model.p);
  A.het = s->het_d.p;
  A.py = s->legacy_py ? s->py.p : nullptr;
  A.U = s->U; A.Z = s->Z; A.K = s->K;
  A.psi = s->psi.p;
  A.path = s->path.p;
  A.nChr = s->nChr; A.bt_rows = s->bt_rows;
  A.tile_first = s->tile_first.p; A.comp = s->bt_comp.p; A.entry = s->bt_entry.p;
  const int K = s->K;
  CK(cudaEventRecord(s->ev[4], s->stream));
  int fP = 0, fPER = 0;
  const bool fast_ok = vit_fast_cfg(K, fP, fPER);
  // 4: warp-reduction scan, 3: cluster-sliced scan, 2: register-sliced dense scan (all one barrier per SNP, emissions
  // materialised first), 0: shared-memory dense scan (any K <= 256, any class matrix)
  const bool cluster_ok = s->U <= 32 && s->Z <= 8;
  const bool redux_ok = s->U <= 64 && s->Z <= 8 && s->vit_monotone;
  int mode = s->vit_mode;
  if (mode < 0) {   // measured on B200 (1 M SNPs, 23 chromosomes)
    if (redux_ok) mode = 4;
    else if (cluster_ok && (s->Z <= kVitClusterAutoMaxZ || !fast_ok)) mode = 3;
    else mode = fast_ok ? 2 : 0;
  }
  if (mode == 4 && !redux_ok) mode = cluster_ok ? 3 : 2;
  if (mode == 3 && !cluster_ok) mode = 2;
  if (mode == 2 && !fast_ok) mode = 0;
  bool dense_from_py = s->legacy_py;
  if (mode >= 2 || s->gauss) {   // (the dense scan evaluates only the binomial emission in place)
    if (!s->legacy_py) {   // exact log emissions, materialised in the (idle) alpha buffer
      const long long total = (long long)s->N * K;
      const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
      viterbi_emission_kernel<<<blocks, 256, 0, s->stream>>>(s->snp_p(), A.M, s->U, K, total, s->alpha.p);
      CK(cudaGetLastError());
      s->tm.kernel_launches += 1;
      s->state_valid = false;
      A.py = s->alpha.p;
      dense_from_py = true;
    }
  }
  if (mode >= 2) {
    if (mode == 4) launch_vit_redux(A, s->nChr, s->stream);
    else if (mode == 3) launch_vit_cluster(A, s->nChr, s->stream);
    else launch_vit_fast(A, s->nChr, K, fP, fPER, s->stream);
  } else {
    launch_vit_dense(A, s->nChr, K, dense_from_py, s->stream);
  }
  CK(cudaGetLastError());
  launch_backtrace(A, s->nTiles, s->stream);
  CK(cudaGetLastError());
  s->tm.kernel_launches += 2 + (s->nTiles > 0 ? 2 : 0);
  CK(cudaEventRecord(s->ev[5], s->stream));
  CK(cudaMemcpyAsync(path_out, s->path.p, sizeof(int) * s->N, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, s->ev[4], s->ev[5]));
  s->tm.last_viterbi_ms = ms;
  return TITAN_OK;
}

extern "C" int titan_viterbi(titan_solver *s, const titan_params *p, int32_t *path_out) {
  if (!s || !p || !path_out) return fail(TITAN_ERR_ARG, "titan_viterbi: NULL argument");
  if (!p->piG || !p->piZ) return fail(TITAN_ERR_ARG, "titan_viterbi: piG and piZ are required");
  try {
    return viterbi_impl(s, p, path_out);
  } catch (const CudaFail &f) {
    return f.code;
  }
}

// ---------------------------------------------------------------------------------------------
// legacy-compatible entries (one chromosome, materialised log emissions)
// ---------------------------------------------------------------------------------------------
// ---- generic compatibility path: any layout the reference's C accepts (outlier state, repeated zygosity
// keys, K not a multiple of numClust, more than 32 unit states) through the dense kernels ----
static bool needs_generic(int K, int numClust, const double *zs, int nZS, int useOutlier) {
  if (useOutlier || numClust <= 0 || K % numClust != 0) return true;
  if (const char *e = getenv("TITAN_FORCE_DENSE")) if (atoi(e)) return true;   // cross-check of the structured kernels
  const int U = K / numClust;
  if (U > 64 || numClust > kMaxZ || nZS < U) return true;
  int nhet = 0;
  for (int a = 0; a < U; ++a) {
    nhet += (zs[a] == -1);
    for (int b = a + 1; b < U; ++b)
      if (zs[a] == zs[b] && zs[a] != -1) return true;
  }
  return nhet > 1;     // several -1 keys form ONE zygosity class in the reference: dense class-matrix path
}

struct GenericSetup {
  int U = 1;
  std::vector<unsigned char> cls;
  std::vector<double> rhoG, rhoZ;
};

static int generic_setup(const char *who, int K, int T, const double *zs, int nZS, int numClust, const double *positions,
                         double zStrength, double txnLen, int useOutlier, GenericSetup *g) {
  if (K <= 0 || T <= 0) return fail(TITAN_ERR_ARG, "%s: The obslik must be %d-by-%d dimension.", who, K, T);
  if (numClust <= 0) return fail(TITAN_ERR_ARG, "%s: numClust must be a positive scalar", who);
  if (!(txnLen > 0) || !(zStrength > 0)) return fail(TITAN_ERR_ARG, "%s: txnLen / zStrength must be positive scalars", who);
  g->U = K / numClust;                                  // src/fwd_backC_clonalCN.c:71 (integer division)
  if (g->U < 1) return fail(TITAN_ERR_ARG, "%s: numClust=%d exceeds the %d states", who, numClust, K);
  const int need = std::min(g->U, useOutlier ? std::max(K - 1, 1) : K);
  if (nZS < need) return fail(TITAN_ERR_ARG, "%s: zygosityKey has %d entries, %d are indexed", who, nZS, need);
  th::generic_class_matrix(K, g->U, zs, useOutlier != 0, g->cls);
  g->rhoG.assign(T, 1.0);
  g->rhoZ.assign(T, 1.0);
  for (int t = 1; t < T; ++t) {
    g->rhoG[t] = th::rho_keep(positions[t - 1], positions[t], txnLen);
    g->rhoZ[t] = th::rho_keep(positions[t - 1], positions[t], zStrength);
  }
  int ndev = 0;
  return titan_device_count(&ndev);
}

static int generic_fwd_back(const double *log_pi, int K, const double *py, int T, const double *zs, int nZS, int numClust,
                            const double *positions, double zStrength, double txnLen, int useOutlier, double *rho_out,
                            double *loglik_out) {
  if (K > 1024) return fail(TITAN_ERR_UNSUPPORTED, "fwd_backC_clonalCN: K=%d exceeds the 1024 states of the dense kernel", K);
  GenericSetup g;
  int rc = generic_setup("fwd_backC_clonalCN", K, T, zs, nZS, numClust, positions, zStrength, txnLen, useOutlier, &g);
  if (rc != TITAN_OK) return rc;
  try {
    DevBuf<double> dpy, dlp, dG, dZ, dal, dlg, dll;
    DevBuf<unsigned char> dcls;
    const size_t KT = (size_t)K * T;
    dpy.alloc(KT); dlp.alloc(K); dG.alloc(T); dZ.alloc(T); dal.alloc(KT); dlg.alloc(KT); dll.alloc(1); dcls.alloc((size_t)K * K);
    CK(cudaMemcpy(dpy.p, py, sizeof(double) * KT, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlp.p, log_pi, sizeof(double) * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dG.p, g.rhoG.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dZ.p, g.rhoZ.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcls.p, g.cls.data(), (size_t)K * K, cudaMemcpyHostToDevice));
    DenseArgs A{dpy.p, dlp.p, dG.p, dZ.p, dcls.p, K, T, dal.p, dlg.p, dll.p};
    const int threads = ((K + 31) / 32) * 32;
    dense_fwd_back_kernel<<<1, threads, sizeof(double) * (K + 32)>>>(A);
    CK(cudaGetLastError());
    CK(cudaMemcpy(rho_out, dlg.p, sizeof(double) * KT, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(loglik_out, dll.p, sizeof(double), cudaMemcpyDeviceToHost));
  } catch (const CudaFail &f) {
    return f.code;
  }
  return TITAN_OK;
}

static int generic_viterbi(const double *log_pi, int K, const double *py, int T, const double *zs, int nZS, int numClust,
                           const double *positions, double zStrength, double txnLen, int useOutlier, int *path_out) {
  if (K > 256) return fail(TITAN_ERR_UNSUPPORTED, "viterbiC_clonalCN: Viterbi back-pointers are 8-bit: K=%d > 256", K);
  GenericSetup g;
  int rc = generic_setup("viterbiC_clonalCN", K, T, zs, nZS, numClust, positions, zStrength, txnLen, useOutlier, &g);
  if (rc != TITAN_OK) return rc;
  try {
    std::unordered_map<std::pair<uint64_t, uint64_t>, int, th::PairHash> ids;
    std::vector<int> id(T, 0);
    std::vector<std::pair<double, double>> distinct;
    for (int t = 1; t < T; ++t) {
      auto key = std::make_pair(th::bits_of(g.rhoG[t]), th::bits_of(g.rhoZ[t]));
      auto it = ids.find(key);
      if (it == ids.end()) {
        it = ids.emplace(key, (int)distinct.size()).first;
        distinct.emplace_back(g.rhoG[t], g.rhoZ[t]);
      }
      id[t] = it->second;
    }
    std::vector<double> tab(std::max<size_t>(1, distinct.size()) * 4 * (size_t)K);
    for (size_t q = 0; q < distinct.size(); ++q)
      th::viterbi_log_table_generic(K, g.cls.data(), distinct[q].first, distinct[q].second, tab.data() + q * 4 * (size_t)K);
    DevBuf<double> dpy, dlp, dtab;
    DevBuf<int> dcs, did, dpath;
    DevBuf<unsigned char> dcls, dpsi, dhet;
    const size_t KT = (size_t)K * T;
    int cs[2] = {0, T};
    std::vector<unsigned char> het0(std::max(g.U, 1), 0);
    dpy.alloc(KT); dlp.alloc(K); dtab.alloc(tab.size()); dcs.alloc(2); did.alloc(T); dpath.alloc(T);
    dcls.alloc((size_t)K * K); dpsi.alloc(KT + 64); dhet.alloc(het0.size());
    int bt_rows = 0;
    const std::vector<int> tf = bt_tile_plan(cs, 1, K, bt_rows);
    DevBuf<int> dtf;
    DevBuf<unsigned char> dcomp, dentry;
    dtf.alloc(tf.size()); dcomp.alloc((size_t)std::max(1, tf.back()) * K); dentry.alloc(std::max(1, tf.back()));
    CK(cudaMemcpy(dtf.p, tf.data(), sizeof(int) * tf.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpy.p, py, sizeof(double) * KT, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlp.p, log_pi, sizeof(double) * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtab.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcs.p, cs, sizeof(cs), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(did.p, id.data(), sizeof(int) * T, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcls.p, g.cls.data(), (size_t)K * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dhet.p, het0.data(), het0.size(), cudaMemcpyHostToDevice));
    VitArgs A{};
    A.chr_start = dcs.p; A.txn_id = did.p; A.ltab = dtab.p; A.het = dhet.p; A.py = dpy.p; A.cls = dcls.p;
    A.M.logpi = dlp.p;
    A.U = g.U; A.Z = numClust; A.K = K; A.psi = dpsi.p; A.path = dpath.p;
    int fP = 0, fPER = 0;
    if (vit_fast_cfg(K, fP, fPER)) launch_vit_fast(A, 1, K, fP, fPER, 0);
    else launch_vit_dense(A, 1, K, true, 0);
    CK(cudaGetLastError());
    A.nChr = 1; A.bt_rows = bt_rows; A.tile_first = dtf.p; A.comp = dcomp.p; A.entry = dentry.p;
    launch_backtrace(A, tf.back(), 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(path_out, dpath.p, sizeof(int) * T, cudaMemcpyDeviceToHost));
  } catch (const CudaFail &f) {
    return f.code;
  }
  return TITAN_OK;
}

static int legacy_create(const char *who, const double *log_pi, int K, const double *py, int T, const double *zs, int nZS,
                         int numClust, const double *positions, double zStrength, double txnLen, int useOutlier,
                         titan_solver **out) {
  if (!log_pi || !py || !zs || !positions) return fail(TITAN_ERR_ARG, "%s: NULL argument", who);
  if (K <= 0 || T <= 0) return fail(TITAN_ERR_ARG, "%s: The obslik must be %d-by-%d dimension.", who, K, T);
  if (numClust <= 0) return fail(TITAN_ERR_ARG, "%s: numClust must be a positive scalar", who);
  if (useOutlier)
    return fail(TITAN_ERR_UNSUPPORTED, "%s: useOutlier=1 is not covered by the structured kernels (no shipped driver "
                                       "enables it; SURVEY.md section 0)", who);
  const int U = K / numClust;   // src/fwd_backC_clonalCN.c:71
  if (U * numClust != K)
    return fail(TITAN_ERR_UNSUPPORTED, "%s: K=%d is not a multiple of numClust=%d", who, K, numClust);
  if (nZS < U) return fail(TITAN_ERR_ARG, "%s: zygosityKey has %d entries, %d unit states", who, nZS, U);
  int64_t cs[2] = {0, T};
  SolverInit in{T, 1, cs, positions, nullptr, nullptr, nullptr, py, txnLen, zStrength, U, numClust, zs, nZS};
  int rc = solver_create_impl(in, out);
  if (rc == TITAN_OK) (*out)->tol = 1e-10;   // compatibility path: tightest setting, it is not the fast path
  return rc;
}

extern "C" int titan_fwd_back_clonalCN(const double *log_piGiZi, int K, const double *py, int T, const double *copyNumKey,
                                       int nCT, const double *zygosityKey, int nZS, int numClust, const double *positions,
                                       double zStrength, double txnLen, int useOutlier, double *rho_out,
                                       double *loglik_out) {
  (void)copyNumKey; (void)nCT;   // unused by the reference's maths as well (fwd_backC_clonalCN.c:266)
  if (!rho_out || !loglik_out) return fail(TITAN_ERR_ARG, "fwd_backC_clonalCN: NULL output");
  if (!log_piGiZi || !py || !zygosityKey || !positions) return fail(TITAN_ERR_ARG, "fwd_backC_clonalCN: NULL argument");
  if (needs_generic(K, numClust, zygosityKey, nZS, useOutlier))
    return generic_fwd_back(log_piGiZi, K, py, T, zygosityKey, nZS, numClust, positions, zStrength, txnLen, useOutlier,
                            rho_out, loglik_out);
  titan_solver *s = nullptr;
  int rc = legacy_create("fwd_backC_clonalCN", log_piGiZi, K, py, T, zygosityKey, nZS, numClust, positions, zStrength,
                         txnLen, useOutlier, &s);
  if (rc != TITAN_OK) return rc;
  try {
    // the initial distribution arrives in log space: pi = exp(log pi)
    ModelConsts H = model_layout(s, s->model_h.p);
    for (int j = 0; j < K; ++j) { ((double *)H.logpi)[j] = log_piGiZi[j]; ((double *)H.pi)[j] = std::exp(log_piGiZi[j]); }
    titan_params p{};
    titan_estep_out o{};
    o.loglik_chr = loglik_out;
    rc = estep_impl(s, &p, &o, false, rho_out);
  } catch (const CudaFail &f) {
    rc = f.code;
  }
  titan_solver_destroy(s);
  return rc;
}

extern "C" int titan_viterbi_clonalCN(const double *log_piGiZi, int K, const double *py, int T, const double *copyNumKey,
                                      int nCT, const double *zygosityKey, int nZS, int numClust, const double *positions,
                                      double zStrength, double txnLen, int useOutlier, int *path_out) {
  (void)copyNumKey; (void)nCT;
  if (!path_out) return fail(TITAN_ERR_ARG, "viterbiC_clonalCN: NULL output");
  if (!log_piGiZi || !py || !zygosityKey || !positions) return fail(TITAN_ERR_ARG, "viterbiC_clonalCN: NULL argument");
  if (needs_generic(K, numClust, zygosityKey, nZS, useOutlier))
    return generic_viterbi(log_piGiZi, K, py, T, zygosityKey, nZS, numClust, positions, zStrength, txnLen, useOutlier,
                           path_out);
  titan_solver *s = nullptr;
  int rc = legacy_create("viterbiC_clonalCN", log_piGiZi, K, py, T, zygosityKey, nZS, numClust, positions, zStrength,
                         txnLen, useOutlier, &s);
  if (rc != TITAN_OK) return rc;
  try {
    ModelConsts H = model_layout(s, s->model_h.p);
    for (int j = 0; j < K; ++j) { ((double *)H.logpi)[j] = log_piGiZi[j]; ((double *)H.pi)[j] = std::exp(log_piGiZi[j]); }
    titan_params p{};
    rc = viterbi_impl(s, &p, path_out);
  } catch (const CudaFail &f) {
    rc = f.code;
  }
  titan_solver_destroy(s);
  return rc;
}

// reference: src/getPositionOverlapC.c:19-50 (host-side; registered but dead in the reference)
extern "C" int titan_get_position_overlap(const double *posn, int T, const double *start, const double *stop, int Tcn,
                                          double *out) {
  if (!posn || !start || !stop || !out) return fail(TITAN_ERR_ARG, "getPositionOverlapC: NULL argument");
  for (int i = 0; i < T; ++i) {
    out[i] = 0;
    for (int j = 0; j < Tcn; ++j)
      if ((int)start[j] <= (int)posn[i] && (int)posn[i] <= (int)stop[j]) { out[i] = j + 1; break; }
  }
  return TITAN_OK;
}

// ---------------------------------------------------------------------------------------------
// M-step and the EM driver
// ---------------------------------------------------------------------------------------------
static int hyper_from_c(const titan_hyper *h, int U, int Z, th::Hyper *H, bool gauss = false) {
  if (!h || !h->rt || !h->ct || !h->var_0 || !h->alphaKHyper || !h->betaKHyper || !h->kappaGHyper || !h->piG_0 ||
      !h->s_0 || !h->alphaSHyper || !h->betaSHyper || !h->kappaZHyper || !h->piZ_0)
    return fail(TITAN_ERR_ARG, "params must include genotypeParams, ploidyParams, normalParams, cellPrevParams.");
  H->U = U; H->Z = Z;
  H->rt.assign(h->rt, h->rt + U); H->ct.assign(h->ct, h->ct + U); H->var_0.assign(h->var_0, h->var_0 + U);
  H->alphaK.assign(h->alphaKHyper, h->alphaKHyper + U); H->betaK.assign(h->betaKHyper, h->betaKHyper + U);
  H->kappaG.assign(h->kappaGHyper, h->kappaGHyper + U); H->piG_0.assign(h->piG_0, h->piG_0 + U);
  H->s_0.assign(h->s_0, h->s_0 + Z); H->alphaS.assign(h->alphaSHyper, h->alphaSHyper + Z);
  H->betaS.assign(h->betaSHyper, h->betaSHyper + Z); H->kappaZ.assign(h->kappaZHyper, h->kappaZHyper + Z);
  H->piZ_0.assign(h->piZ_0, h->piZ_0 + Z);
  H->rn = h->rn; H->n_0 = h->n_0; H->alphaN = h->alphaNHyper; H->betaN = h->betaNHyper;
  H->phi_0 = h->phi_0; H->alphaP = h->alphaPHyper; H->betaP = h->betaPHyper;
  H->gauss = gauss;
  if (gauss) {
    if (!h->varR_0 || !h->alphaRHyper || !h->betaRHyper)
      return fail(TITAN_ERR_ARG, "genotypeParams must contain varR_0, alphaRHyper, betaRHyper and corRho_0 for the Gaussian "
                                 "allele model. See loadDefaultParameters(data=).");
    if (!(h->corRho_0 > -1 && h->corRho_0 < 1)) return fail(TITAN_ERR_ARG, "genotypeParams$corRho_0=%g outside (-1, 1)", h->corRho_0);
    H->varR_0.assign(h->varR_0, h->varR_0 + U); H->alphaR.assign(h->alphaRHyper, h->alphaRHyper + U);
    H->betaR.assign(h->betaRHyper, h->betaRHyper + U);
    H->corRho = h->corRho_0;
  }
  return TITAN_OK;
}

// estimateClonalCNParamsMap, R/paramEstimation.R:10-70 given the six sums and the posteriors at chrPos_0
static int mstep_core(const th::Hyper &H, int nChr, int maxiterUpdate, bool normal_map, bool estS, bool estP,
                      const double *stats, const double *rhoG0, const double *rhoZ0, double prev_n, const double *prev_s,
                      const double *prev_var, double prev_phi, const double *prev_piG, const double *prev_piZ,
                      double *out_n, double *out_s, double *out_var, double *out_phi, double *out_piG, double *out_piZ,
                      int *out_sweeps, const double *prev_varR = nullptr, double *out_varR = nullptr) {
  const int U = H.U, Z = H.Z, K = U * Z;
  th::Stats st{stats, stats + K, stats + 2 * K, stats + 3 * K, stats + 4 * K, stats + 5 * K};
  // Gaussian allele model: the six slots hold a, f, c, h, e, g (titan_common.cuh)
  th::GStats gst{stats, stats + 2 * K, stats + 4 * K, stats + K, stats + 5 * K, stats + 3 * K};
  // likInitG/Z: sum(t(log(pi_0)) %*% rho[, chrPos_0])  (:43-44)
  double likG = 0, likZ = 0, totG = 0, totZ = 0;
  for (int c = 0; c < nChr; ++c) {
    double acc = 0;
    for (int u = 0; u < U; ++u) { acc += std::log(prev_piG[u]) * rhoG0[(size_t)c * U + u]; totG += rhoG0[(size_t)c * U + u]; }
    likG += acc;
    acc = 0;
    for (int z = 0; z < Z; ++z) { acc += std::log(prev_piZ[z]) * rhoZ0[(size_t)c * Z + z]; totZ += rhoZ0[(size_t)c * Z + z]; }
    likZ += acc;
  }
  th::MStepOut mo = H.gauss ? th::update_parameters_gauss(H, gst, prev_n, prev_s, prev_var, prev_varR, prev_phi, prev_piG,
                                                          prev_piZ, likG, likZ, normal_map, estS, estP, maxiterUpdate, 10)
                            : th::update_parameters(H, st, prev_n, prev_s, prev_var, prev_phi, prev_piG, prev_piZ, likG, likZ,
                                                    normal_map, estS, estP, maxiterUpdate, 10);
  *out_n = mo.n; *out_phi = mo.phi;
  std::copy(mo.s.begin(), mo.s.end(), out_s);
  std::copy(mo.var.begin(), mo.var.end(), out_var);
  if (H.gauss && out_varR) std::copy(mo.varR.begin(), mo.varR.end(), out_varR);
  // pi updates: sum() over the whole sub-matrix (R/paramEstimation.R:533-547)
  double skZ = 0, skG = 0;
  for (int z = 0; z < Z; ++z) skZ += H.kappaZ[z];
  for (int u = 0; u < U; ++u) skG += H.kappaG[u];
  for (int z = 0; z < Z; ++z) out_piZ[z] = (totZ + H.kappaZ[z] - 1) / (totZ + skZ - Z);
  for (int u = 0; u < U; ++u) out_piG[u] = (totG + H.kappaG[u] - 1) / (totG + skG - U);
  if (out_sweeps) *out_sweeps = mo.sweeps;
  return TITAN_OK;
}

extern "C" int titan_mstep(const titan_hyper *h, int U, int Z, int nChr, int maxiterUpdate, int normal_map, int estimateS,
                           int estimatePloidy, const double *stats, const double *rhoG0, const double *rhoZ0, double prev_n,
                           const double *prev_s, const double *prev_var, double prev_phi, const double *prev_piG,
                           const double *prev_piZ, double *out_n, double *out_s, double *out_var, double *out_phi,
                           double *out_piG, double *out_piZ, int *out_sweeps, double *out_prior) {
  th::Hyper H;
  int rc = hyper_from_c(h, U, Z, &H);
  if (rc != TITAN_OK) return rc;
  if (!stats || !rhoG0 || !rhoZ0 || !prev_s || !prev_var || !prev_piG || !prev_piZ || !out_n || !out_s || !out_var ||
      !out_phi || !out_piG || !out_piZ)
    return fail(TITAN_ERR_ARG, "titan_mstep: NULL argument");
  rc = mstep_core(H, nChr, maxiterUpdate, normal_map != 0, estimateS != 0, estimatePloidy != 0, stats, rhoG0, rhoZ0, prev_n,
                  prev_s, prev_var, prev_phi, prev_piG, prev_piZ, out_n, out_s, out_var, out_phi, out_piG, out_piZ,
                  out_sweeps);
  if (rc == TITAN_OK && out_prior) {
    th::Priors pr = th::prior_probs(H, *out_n, out_s, *out_phi, out_var, out_piG, out_piZ, normal_map != 0, estimateS != 0,
                                    estimatePloidy != 0);
    *out_prior = pr.prior;
  }
  return rc;
}

extern "C" int titan_mstep_gaussian(const titan_hyper *h, int U, int Z, int nChr, int maxiterUpdate, int normal_map,
                                    int estimateS, int estimatePloidy, const double *stats, const double *rhoG0,
                                    const double *rhoZ0, double prev_n, const double *prev_s, const double *prev_var,
                                    const double *prev_varR, double prev_phi, const double *prev_piG, const double *prev_piZ,
                                    double *out_n, double *out_s, double *out_var, double *out_varR, double *out_phi,
                                    double *out_piG, double *out_piZ, int *out_sweeps, double *out_prior) {
  th::Hyper H;
  int rc = hyper_from_c(h, U, Z, &H, true);
  if (rc != TITAN_OK) return rc;
  if (!stats || !rhoG0 || !rhoZ0 || !prev_s || !prev_var || !prev_varR || !prev_piG || !prev_piZ || !out_n || !out_s ||
      !out_var || !out_varR || !out_phi || !out_piG || !out_piZ)
    return fail(TITAN_ERR_ARG, "titan_mstep_gaussian: NULL argument");
  rc = mstep_core(H, nChr, maxiterUpdate, normal_map != 0, estimateS != 0, estimatePloidy != 0, stats, rhoG0, rhoZ0, prev_n,
                  prev_s, prev_var, prev_phi, prev_piG, prev_piZ, out_n, out_s, out_var, out_phi, out_piG, out_piZ,
                  out_sweeps, prev_varR, out_varR);
  if (rc == TITAN_OK && out_prior) {
    th::Priors pr = th::prior_probs(H, *out_n, out_s, *out_phi, out_var, out_piG, out_piZ, normal_map != 0, estimateS != 0,
                                    estimatePloidy != 0, out_varR);
    *out_prior = pr.prior;
  }
  return rc;
}

// runEMclonalCN, R/hmmClonal.R:8-368 (either allele model, useOutlierState = FALSE)
extern "C" int titan_em_run(titan_solver *s, const titan_hyper *h, const titan_em_opts *o, titan_em_out *out) {
  if (!s || !h || !o || !out) return fail(TITAN_ERR_ARG, "titan_em_run: NULL argument");
  if (s->legacy_py) return fail(TITAN_ERR_ARG, "titan_em_run: solver was created for materialised emissions");
  const int U = s->U, Z = s->Z, K = s->K, nChr = s->nChrOut();      // whole genome when the solution is sharded
  th::Hyper H;
  int rc = hyper_from_c(h, U, Z, &H, s->gauss);
  if (rc != TITAN_OK) return rc;
  if (!out->n || !out->phi || !out->loglik || !out->s || !out->piZ || !out->var || !out->piG || !out->muR || !out->muC)
    return fail(TITAN_ERR_ARG, "titan_em_run: output arrays are required");
  // allelic variances per iteration (:100,:132,:252); the caller's array when given
  std::vector<double> varR_own;
  double *varR = out->varR;
  if (s->gauss && !varR) { varR_own.assign((size_t)U * (o->maxiter + 1), 0.0); varR = varR_own.data(); }
  const int maxit = o->maxiter + 1;   // :92
  const bool nmap = o->normal_map != 0, estS = o->estimateS != 0, estP = o->estimatePloidy != 0;
  double *n = out->n, *phi = out->phi, *loglik = out->loglik;
  // The N-sized posterior outputs are written once, after the last iteration.  Freshly allocated host arrays
  // (R's allocMatrix, numpy's zeros) have no pages behind them yet, and faulting 8(U+Z)N bytes in during the
  // final device-to-host copy costs more than the copy: touch them on a worker thread while the EM runs.
  struct Prefault {
    std::thread t;
    ~Prefault() { if (t.joinable()) t.join(); }
  } prefault;
  if (o->want_posteriors && (out->rhoG || out->rhoZ)) {
    double *pg = out->rhoG, *pz = out->rhoZ;
    const size_t ng = (size_t)s->N * U, nz = (size_t)s->N * Z;
    prefault.t = std::thread([pg, pz, ng, nz]() {
      if (pg) std::memset(pg, 0, ng * sizeof(double));
      if (pz) std::memset(pz, 0, nz * sizeof(double));
    });
  }
  auto col = [](double *base, int rows, int i) { return base + (size_t)rows * i; };
  int i = 0;
  loglik[0] = -std::numeric_limits<double>::infinity();
  std::copy(H.s_0.begin(), H.s_0.end(), col(out->s, Z, 0));
  std::copy(H.var_0.begin(), H.var_0.end(), col(out->var, U, 0));
  phi[0] = H.phi_0;
  n[0] = H.n_0;
  std::copy(H.piG_0.begin(), H.piG_0.end(), col(out->piG, U, 0));
  std::copy(H.piZ_0.begin(), H.piZ_0.end(), col(out->piZ, Z, 0));
  if (varR) {
    if (s->gauss) std::copy(H.varR_0.begin(), H.varR_0.end(), col(varR, U, 0));
    else if (h->varR_0) std::copy(h->varR_0, h->varR_0 + U, col(varR, U, 0));
    else std::fill(col(varR, U, 0), col(varR, U, 0) + U, 1.0 / 20);
  }
  th::mixture_means(U, Z, H.rt.data(), H.rn, n[0], col(out->s, Z, 0), H.ct.data(), phi[0], col(out->muR, K, 0),
                    col(out->muC, K, 0));
  std::vector<double> stats(6 * (size_t)K), llc(nChr), rG0((size_t)nChr * U), rZ0((size_t)nChr * Z);
  bool converged = false;
  out->estep_ms = 0;
  out->mstep_ms = 0;
  out->reduce_ms = 0;
  titan_params last_p{};
  bool have_last = false;
  try {
    while (!converged && (i + 1) < maxit) {
      ++i;
      titan_params p{};
      p.n = n[i - 1]; p.phi = phi[i - 1]; p.s = col(out->s, Z, i - 1); p.var = col(out->var, U, i - 1);
      p.piG = col(out->piG, U, i - 1); p.piZ = col(out->piZ, Z, i - 1);
      p.rt = H.rt.data(); p.ct = H.ct.data(); p.rn = H.rn;
      if (s->gauss) { p.varR = col(varR, U, i - 1); p.corRho = H.corRho; }
      titan_estep_out eo{};
      eo.loglik_chr = llc.data(); eo.stats = stats.data(); eo.rhoG0 = rG0.data(); eo.rhoZ0 = rZ0.data();
      rc = estep_impl(s, &p, &eo, false, nullptr);
      if (rc != TITAN_OK) return rc;
      last_p = p;
      have_last = true;
      out->estep_ms += s->tm.last_estep_ms;
      out->reduce_ms += s->tm.last_reduce_ms;
      double ll = 0;
      for (int c = 0; c < nChr; ++c) ll += llc[c];        // :218 sum over chromosomes
      if (!std::isfinite(ll)) return fail(TITAN_ERR_NUMERIC, "non-finite log-likelihood in EM iteration %d", i);
      auto t0 = std::chrono::steady_clock::now();
      int sweeps = 0;
      rc = mstep_core(H, nChr, o->maxiterUpdate, nmap, estS, estP, stats.data(), rG0.data(), rZ0.data(), n[i - 1],
                      col(out->s, Z, i - 1), col(out->var, U, i - 1), phi[i - 1], col(out->piG, U, i - 1),
                      col(out->piZ, Z, i - 1), &n[i], col(out->s, Z, i), col(out->var, U, i), &phi[i], col(out->piG, U, i),
                      col(out->piZ, Z, i), &sweeps, s->gauss ? col(varR, U, i - 1) : nullptr,
                      s->gauss ? col(varR, U, i) : nullptr);
      if (rc != TITAN_OK) return rc;
      if (varR && !s->gauss) std::copy(col(varR, U, i - 1), col(varR, U, i - 1) + U, col(varR, U, i));   // :212
      if (out->cd_sweeps) out->cd_sweeps[i] = sweeps;
      th::mixture_means(U, Z, H.rt.data(), H.rn, n[i], col(out->s, Z, i), H.ct.data(), phi[i], col(out->muR, K, i),
                        col(out->muC, K, i));
      th::Priors pr = th::prior_probs(H, n[i], col(out->s, Z, i), phi[i], col(out->var, U, i), col(out->piG, U, i),
                                      col(out->piZ, Z, i), nmap, estS, estP, s->gauss ? col(varR, U, i) : nullptr);
      loglik[i] = ll + pr.prior;                          // :311
      out->mstep_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (o->verbose)
        fprintf(stderr, "fwdBack: EM iteration %d complete loglik=%0.4f n=%0.4f phi=%0.4f (cd sweeps %d, fwd/bwd rounds %d/%d)\n",
                i, loglik[i], n[i], phi[i], sweeps, s->tm.last_fwd_rounds, s->tm.last_bwd_rounds);
      if ((std::fabs(loglik[i] - loglik[i - 1]) / std::fabs(loglik[i])) <= o->likChangeThreshold && loglik[i] >= loglik[i - 1]) {
        converged = true;                                 // :316-318
      } else if (loglik[i] < loglik[i - 1]) {
        converged = true;                                 // :319-325 roll back one column
        --i;
      }
    }
    out->iters = i + 1;
    // rhoG / rhoZ of the last executed E-step (R/hmmClonal.R:356-357): the per-iteration passes do not
    // write the N-sized marginals; its forward variables and verified boundaries are still resident,
    // so only the final backward pass is repeated with the output switched on.
    if (prefault.t.joinable()) prefault.t.join();
    if (o->want_posteriors && have_last && (out->rhoG || out->rhoZ)) {
      titan_estep_out eo{};
      eo.rhoG = out->rhoG;
      eo.rhoZ = out->rhoZ;
      rc = estep_impl(s, &last_p, &eo, true, nullptr, /*posterior_only=*/true);
      if (rc != TITAN_OK) return rc;
      out->estep_ms += s->tm.last_estep_ms;
    }
  } catch (const CudaFail &f) {
    return f.code;
  }
  return TITAN_OK;
}
