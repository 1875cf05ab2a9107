// titan_kernels.cuh -- sm_100a kernels of the TitanCNA E-step hot path (fp64 throughout).
//
// State space: joint state j = u + z*U (genotype u fastest; reference R/hmmClonal.R:196,
// src/fwd_backC_clonalCN.c:234-235).  One WARP owns one position chunk of one chromosome:
// lane g holds the Z cluster values of genotype g in registers, so the per-genotype sum is
// thread-local and the per-cluster sums are Z butterfly reductions over the warp.
//
// Transition structure (src/fwd_backC_clonalCN.c:215-301): raw[i,j] is one of four values,
// selected by (same genotype?, same cluster or destination is diploid-HET?), rows divided by
// their sum.  Regrouped so every coefficient is non-negative (no cancellation):
//   destination non-HET:  o[g,z] = k0*W + k1*Wz[z] + k2*Wg[g] + k3*w[g,z]
//   destination HET    :  o[g,z] = cB*W + k4*Wg[g]
// with w = alpha/rowsum(source), k0 = oG(1-rhoZ), k1 = oG(2rhoZ-1), k2 = (rhoG-oG)(1-rhoZ),
// k3 = (rhoG-oG)(2rhoZ-1), cB = oG*rhoZ, k4 = (rhoG-oG)rhoZ, oG = (1-rhoG)/(K-1).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace titan {

struct __align__(16) SnpRec {   // parameter-independent per-SNP columns (HBM, read twice/E-step)
  double x;    // reference count              (data$ref)
  double nx;   // tumDepth - ref
  double lr;   // logR (natural log)
  double lb;   // lgamma(N+1)-lgamma(x+1)-lgamma(N-x+1), host libm (R/hmmClonal.R:618)
};

struct __align__(16) TxnRec {   // structured transition between SNP t-1 and t (unused at chr start)
  double k0, k1, k2, k3, cB, k4, irn, irh;   // irn/irh = 1/rowsum of a non-HET / HET source
};

struct Chunk {
  int t0, t1;      // SNP range [t0, t1) owned by this chunk
  int cs, ce;      // chromosome range [cs, ce)
  int chr;         // chromosome index
  int pad;
};

struct ModelConsts {            // device pointers, refreshed every E-step (K- and U-sized)
  const double *lmu;    // [K] log(muR_j)
  const double *l1mu;   // [K] log(1-muR_j)
  const double *muC;    // [K]
  const double *cN;     // [U] log(1/(sqrt(var_u)*sqrt(2*pi)))
  const double *i2v;    // [U] 1/(2*var_u)       (forward-backward)
  const double *tv;     // [U] 2*var_u           (Viterbi, exact division as in the reference)
  const double *pi;     // [K] piG_u*piZ_z
  const double *logpi;  // [K] host-libm log of the above
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 2^-exponent(v) for positive finite v (exact power of two; 1.0 for 0 / non-finite input)
__device__ __forceinline__ double pow2_rescale(double v, int &e_out) {
  int hi = __double2hiint(v);
  int be = (hi >> 20) & 0x7ff;
  if (be == 0 || be == 0x7ff) { e_out = 0; return 1.0; }
  e_out = be - 1023;
  return __hiloint2double((2046 - be) << 20, 0);   // biased exponent 1023 - e
}

// ------------------------------------------------------------------------------------------
// emission: B[z] = exp(py[g,z,t] - m_t), m_t ~ max_j py[j,t] rounded to float (any common shift
// is exact as long as it is accounted for in the log-likelihood).
// reference: R/hmmClonal.R:617-634 (densities), :491-523 (state loops)
// ------------------------------------------------------------------------------------------
template <int Z>
struct LaneModel {
  double lmu[Z], l1mu[Z], muC[Z];
  double cN, i2v;
  bool active, het;
};

template <int Z>
__device__ __forceinline__ void load_lane_model(LaneModel<Z> &L, const ModelConsts &M, int U, int lane,
                                                const unsigned char *het_flags) {
  L.active = lane < U;
  L.het = L.active && het_flags[lane];
  int g = L.active ? lane : 0;
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    L.lmu[z] = M.lmu[g + z * U];
    L.l1mu[z] = M.l1mu[g + z * U];
    L.muC[z] = M.muC[g + z * U];
  }
  L.cN = M.cN[g];
  L.i2v = M.i2v[g];
}

template <int Z>
__device__ __forceinline__ double emission(const LaneModel<Z> &L, const SnpRec &r, double (&B)[Z]) {
  double py[Z];
  double mx = -CUDART_INF;
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    double d = r.lr - L.muC[z];
    double v = (r.lb + (r.x * L.lmu[z] + r.nx * L.l1mu[z])) + (L.cN - d * d * L.i2v);
    py[z] = L.active ? v : -CUDART_INF;
    mx = fmax(mx, py[z]);
  }
  float mf = warp_maxf(__double2float_rn(mx));
  double m = (double)mf;
#pragma unroll
  for (int z = 0; z < Z; ++z) B[z] = L.active ? exp(py[z] - m) : 0.0;
  return m;
}

// same, log-emissions read from a materialised K x T matrix (legacy .Call boundary)
template <int Z>
__device__ __forceinline__ double emission_from_py(const double *__restrict__ py_t, int U, int lane, bool active,
                                                   double (&B)[Z]) {
  double py[Z];
  double mx = -CUDART_INF;
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    py[z] = active ? __ldg(py_t + lane + z * U) : -CUDART_INF;
    mx = fmax(mx, py[z]);
  }
  float mf = warp_maxf(__double2float_rn(mx));
  // a float-rounded shift can sit below the true max by one float ulp; harmless
  double m = (mf == -CUDART_INF_F) ? 0.0 : (double)mf;
#pragma unroll
  for (int z = 0; z < Z; ++z) B[z] = active ? exp(py[z] - m) : 0.0;
  return m;
}

struct FbArgs {
  const SnpRec *snp;          // [N]
  const TxnRec *txn;          // [N]
  const Chunk *chunks;        // [nChunks]
  int nChunks;
  int U;
  int warm;                   // warm-up length (SNPs) for speculative chunk starts
  int round;                  // 0 = first pass; >0 = verification / re-run pass
  double rtol, atol;
  ModelConsts M;
  const unsigned char *het;   // [U]
  const double *py;           // [K x N] materialised log emissions (legacy mode) or nullptr
  double *alpha;              // [N][K] scaled forward variables
  double *startv;             // [nChunks][K] boundary vector each chunk last started from
  double *endv;               // [2][nChunks][K] boundary vector each chunk produced (double-buffered by round)
  double *chunk_ll;           // [nChunks] log-likelihood piece of each chunk
  unsigned int *rerun;        // [max_rounds] number of chunks (re)run in each round
  // backward / posterior outputs
  double *part;               // [nChunks][6][K] per-chunk statistic partial sums
  double *rhoG0, *rhoZ0;      // [nChr][U], [nChr][Z]
  double *rhoG, *rhoZ;        // optional [N][U], [N][Z]
  double *loggamma;           // optional [N][K] (legacy mode)
};

template <int Z>
__device__ __forceinline__ bool vec_mismatch(const double (&a)[Z], const double (&b)[Z], double rtol, double atol) {
  bool bad = false;
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    double hi = fmax(a[z], b[z]);
    bad |= (fabs(a[z] - b[z]) > rtol * hi) && (hi > atol);
    bad |= (a[z] != a[z]) || (b[z] != b[z]);
  }
  return __any_sync(0xffffffffu, bad);
}

// ------------------------------------------------------------------------------------------
// Record streaming.  The recursion is a dependent chain, so every cycle of load latency inside
// a step is paid 10^5 times per chromosome.  Each warp therefore stages its per-SNP records
// (SnpRec 32 B + TxnRec 64 B) through a private shared-memory ring of two 16-step tiles: a
// tile is fetched from HBM with coalesced 16-byte loads one-and-a-half tiles ahead of its use,
// parked in registers, and dropped into the ring when the tile before it retires.  Steps then
// read their records with broadcast shared-memory loads one step ahead of the recursion.
// DIR = +1 walks t upwards (forward pass), -1 downwards (backward pass).
// ------------------------------------------------------------------------------------------
constexpr int kTile = 16;
constexpr int kTileBytes = kTile * (32 + 64);          // 1536
constexpr int kWarpSmem = 2 * kTileBytes;              // 3072 B per warp

template <int DIR>
struct RecStream {
  const SnpRec *snp;
  const TxnRec *txn;
  char *sm;
  int tbeg, lo, hi, lane;
  uint4 r0, r1, r2;

  __device__ __forceinline__ int gidx(int q) const { return min(max(tbeg + DIR * q, lo), hi); }
  __device__ __forceinline__ void gload(int k) {
    r0 = __ldg(reinterpret_cast<const uint4 *>(snp + gidx(kTile * k + (lane >> 1))) + (lane & 1));
    r1 = __ldg(reinterpret_cast<const uint4 *>(txn + gidx(kTile * k + (lane >> 2))) + (lane & 3));
    r2 = __ldg(reinterpret_cast<const uint4 *>(txn + gidx(kTile * k + 8 + (lane >> 2))) + (lane & 3));
  }
  __device__ __forceinline__ void sstore(int k) {
    char *b = sm + (k & 1) * kTileBytes;
    reinterpret_cast<uint4 *>(b)[lane] = r0;
    reinterpret_cast<uint4 *>(b + kTile * 32)[lane] = r1;
    reinterpret_cast<uint4 *>(b + kTile * 32)[32 + lane] = r2;
  }
  __device__ __forceinline__ void init(const SnpRec *s, const TxnRec *t, char *smem, int tbeg_, int lo_, int hi_, int lane_) {
    snp = s; txn = t; sm = smem; tbeg = tbeg_; lo = lo_; hi = hi_; lane = lane_;
    gload(0); sstore(0);
    gload(1); sstore(1);
    gload(2);
    __syncwarp();
  }
  // call at the top of step q: when a new tile starts, retire the previous one
  __device__ __forceinline__ void advance(int q) {
    if ((q & (kTile - 1)) == 0 && q > 0) {
      const int k = q / kTile;
      __syncwarp();
      sstore(k + 1);
      gload(k + 2);
      __syncwarp();
    }
  }
  __device__ __forceinline__ void get(int q, SnpRec &s, TxnRec &t) const {
    const char *b = sm + ((q / kTile) & 1) * kTileBytes;
    const int i = q & (kTile - 1);
    s = *reinterpret_cast<const SnpRec *>(b + 32 * i);
    t = *reinterpret_cast<const TxnRec *>(b + kTile * 32 + 64 * i);
  }
};

// one structured forward step: a <- (A_t^T applied to a) .* (B * 2^-e)
template <int Z>
__device__ __forceinline__ void fwd_step(double (&a)[Z], const double (&B)[Z], const TxnRec &tr, bool het, int &e) {
  const double irs = het ? tr.irh : tr.irn;
  double w[Z], Wz[Z];
  double Wg = 0.0;
#pragma unroll
  for (int z = 0; z < Z; ++z) { w[z] = a[z] * irs; Wg += w[z]; }
#pragma unroll
  for (int z = 0; z < Z; ++z) Wz[z] = warp_sum(w[z]);
  double W = 0.0;
#pragma unroll
  for (int z = 0; z < Z; ++z) W += Wz[z];
  const double sc = pow2_rescale(W, e);
  const double base = het ? (tr.cB * W + tr.k4 * Wg) : (tr.k0 * W + tr.k2 * Wg);
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    const double o = het ? base : (base + tr.k1 * Wz[z] + tr.k3 * w[z]);
    a[z] = o * (B[z] * sc);
  }
}

// one structured backward step: b <- A_{t+1} applied to (b .* B_{t+1}), rescaled by a power of two
template <int Z>
__device__ __forceinline__ void bwd_step(double (&b)[Z], const double (&B)[Z], const TxnRec &tr, bool het, bool active) {
  double bb[Z], Wzn[Z];
  double loc = 0.0;
#pragma unroll
  for (int z = 0; z < Z; ++z) { bb[z] = b[z] * B[z]; loc += bb[z]; }
#pragma unroll
  for (int z = 0; z < Z; ++z) Wzn[z] = warp_sum(het ? 0.0 : bb[z]);
  const double H = warp_sum(het ? loc : 0.0);
  double Wn = 0.0;
#pragma unroll
  for (int z = 0; z < Z; ++z) Wn += Wzn[z];
  int e;
  const double sc = pow2_rescale(Wn + H, e);
  const double irs = (het ? tr.irh : tr.irn) * sc;
  const double base = tr.k0 * Wn + tr.cB * H + (het ? tr.k4 * loc : tr.k2 * loc);
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    const double o = base + tr.k1 * Wzn[z] + (het ? 0.0 : tr.k3 * bb[z]);
    b[z] = active ? o * irs : 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// Forward pass over one chunk per warp.
// reference recursion: src/fwd_backC_clonalCN.c:111-142 (log space, dense) -- here scaled linear
// space with the structured O(K) step; log-likelihood pieces are exact for any scaling.
// The loop body holds two independent instruction streams -- the recursion for SNP t and the
// emission (exp) for SNP t+1 -- so that their latencies overlap.
// ------------------------------------------------------------------------------------------
template <int Z, bool FROM_PY>
__global__ void __launch_bounds__(128) fwd_chunk_kernel(FbArgs A) {
  __shared__ __align__(16) char smem_all[4 * kWarpSmem];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= A.nChunks) return;
  const Chunk c = A.chunks[warp];
  const int U = A.U, K = U * Z;
  const bool active = lane < U;
  const size_t cur = (size_t)(A.round & 1) * A.nChunks * K, prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  double *endv_cur = A.endv + cur + (size_t)warp * K;
  const double *endv_prv_self = A.endv + prv + (size_t)warp * K;
  double *startv = A.startv + (size_t)warp * K;

  double a[Z];
  int tb;               // first SNP index processed
  bool have_start;      // a[] already holds alpha~_{tb-1}
  if (A.round > 0) {
    if (c.t0 == c.cs) {   // chromosome-leading chunks start from pi: exact in round 0
      if (active)
#pragma unroll
        for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = endv_prv_self[lane + z * U];
      return;
    }
    double used[Z], want[Z];
    const double *pe = A.endv + prv + (size_t)(warp - 1) * K;
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      used[z] = active ? startv[lane + z * U] : 0.0;
      want[z] = active ? pe[lane + z * U] : 0.0;
    }
    if (!vec_mismatch<Z>(used, want, A.rtol, A.atol)) {
      if (active)
#pragma unroll
        for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = endv_prv_self[lane + z * U];
      return;
    }
#pragma unroll
    for (int z = 0; z < Z; ++z) a[z] = want[z];
    tb = c.t0;
    have_start = true;
  } else {
    tb = max(c.t0 - A.warm, c.cs);
    have_start = false;
#pragma unroll
    for (int z = 0; z < Z; ++z) a[z] = 0.0;
  }
  if (lane == 0) atomicAdd(A.rerun + A.round, 1u);

  LaneModel<Z> L;
  if (!FROM_PY) load_lane_model<Z>(L, A.M, U, lane, A.het);
  const bool het = active && A.het[lane];

  RecStream<1> rs;
  rs.init(A.snp, A.txn, smem_all + (threadIdx.x >> 5) * kWarpSmem, tb, c.cs, c.ce - 1, lane);
  const int n = c.t1 - tb;           // steps q = 0..n-1, t = tb + q
  const int q0 = c.t0 - tb;          // first owned step

  SnpRec rec;
  TxnRec tr;
  double B[Z];
  double m;
  rs.get(0, rec, tr);
  if (FROM_PY) m = emission_from_py<Z>(A.py + (size_t)tb * K, U, lane, active, B);
  else m = emission<Z>(L, rec, B);

  double acc_m = 0.0;     // sum of emission shifts over the owned range
  long long acc_e = 0;    // sum of power-of-two rescale exponents over the owned range
  double s_start = 1.0;   // mass of the vector entering the owned range (1 for a chromosome-leading chunk)
  int q = 0;
  // ---- step q = 0 without a predecessor vector (chromosome start or speculative flat start) ----
  if (!have_start) {
    if (tb == c.cs) {
#pragma unroll
      for (int z = 0; z < Z; ++z) a[z] = (active ? A.M.pi[lane + z * U] : 0.0) * B[z];
    } else {
#pragma unroll
      for (int z = 0; z < Z; ++z) a[z] = B[z];
    }
    if (q0 == 0) {        // only a chromosome-leading chunk owns its very first step
      acc_m += m;
      if (active)
#pragma unroll
        for (int z = 0; z < Z; ++z) A.alpha[(size_t)tb * K + lane + z * U] = a[z];
    }
    q = 1;
    if (n > 1) {
      rs.get(1, rec, tr);
      if (FROM_PY) m = emission_from_py<Z>(A.py + (size_t)(tb + 1) * K, U, lane, active, B);
      else m = emission<Z>(L, rec, B);
    }
  }
  // ---- warm-up steps (not owned, nothing stored) ----
  for (; q < q0; ++q) {
    rs.advance(q);
    SnpRec recn;
    TxnRec trn;
    rs.get(min(q + 1, n - 1), recn, trn);
    double Bn[Z];
    double mn;
    if (FROM_PY) mn = emission_from_py<Z>(A.py + (size_t)min(tb + q + 1, c.t1 - 1) * K, U, lane, active, Bn);
    else mn = emission<Z>(L, recn, Bn);
    int e;
    fwd_step<Z>(a, B, tr, het, e);
    tr = trn; m = mn;
#pragma unroll
    for (int z = 0; z < Z; ++z) B[z] = Bn[z];
  }
  // ---- entering the owned range: record / publish the boundary vector alpha~_{t0-1} ----
  if (c.t0 != c.cs) {
    double loc = 0.0;
#pragma unroll
    for (int z = 0; z < Z; ++z) loc += a[z];
    s_start = warp_sum(loc);
    if (!have_start) {
      const double inv = 1.0 / s_start;
#pragma unroll
      for (int z = 0; z < Z; ++z) a[z] *= inv;
      s_start = 1.0;
    }
    if (active)
#pragma unroll
      for (int z = 0; z < Z; ++z) startv[lane + z * U] = a[z];
  }
  // ---- owned steps ----
  for (; q < n; ++q) {
    rs.advance(q);
    SnpRec recn;
    TxnRec trn;
    rs.get(min(q + 1, n - 1), recn, trn);
    double Bn[Z];
    double mn;
    if (FROM_PY) mn = emission_from_py<Z>(A.py + (size_t)min(tb + q + 1, c.t1 - 1) * K, U, lane, active, Bn);
    else mn = emission<Z>(L, recn, Bn);
    int e;
    fwd_step<Z>(a, B, tr, het, e);
    acc_m += m;
    acc_e += e;
    if (active) {
      double *dst = A.alpha + (size_t)(tb + q) * K + lane;
#pragma unroll
      for (int z = 0; z < Z; ++z) dst[z * U] = a[z];
    }
    tr = trn; m = mn;
#pragma unroll
    for (int z = 0; z < Z; ++z) B[z] = Bn[z];
  }
  // chunk end: boundary vector (normalised) + log-likelihood piece
  double loc = 0.0;
#pragma unroll
  for (int z = 0; z < Z; ++z) loc += a[z];
  const double s_end = warp_sum(loc);
  const double inv = 1.0 / s_end;
  if (active) {
#pragma unroll
    for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = a[z] * inv;
  }
  if (lane == 0)
    A.chunk_ll[warp] = (log(s_end) - log(s_start)) + acc_m + (double)acc_e * 0.6931471805599453094;
}

// ------------------------------------------------------------------------------------------
// Backward pass + posteriors + sufficient statistics over one chunk per warp.
// reference: src/fwd_backC_clonalCN.c:148-192 (beta, gamma); R/hmmClonal.R:230-234 (exp, marginals);
// R/paramEstimation.R:34-40 (a,b,c,d,e,g)
// Iteration for SNP t (descending): beta~_t from beta~_{t+1} and B_{t+1} (carried), posterior at t,
// and -- independent of both -- the emission B_t for the next iteration.
// ------------------------------------------------------------------------------------------
template <int Z, bool FROM_PY, bool WRITE_POST>
__global__ void __launch_bounds__(128) bwd_chunk_kernel(FbArgs A) {
  __shared__ __align__(16) char smem_all[4 * kWarpSmem];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= A.nChunks) return;
  const Chunk c = A.chunks[warp];
  const int U = A.U, K = U * Z;
  const bool active = lane < U;
  const size_t cur = (size_t)(A.round & 1) * A.nChunks * K, prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  double *endv_cur = A.endv + cur + (size_t)warp * K;
  const double *endv_prv_self = A.endv + prv + (size_t)warp * K;
  double *startv = A.startv + (size_t)warp * K;

  // b[] holds beta~_{tn}; tn == c.ce means "past the end" (beta_{ce-1} = 1)
  double b[Z];
  int tn;
  bool from_boundary;   // b[] was read from the next chunk's published boundary vector
  if (A.round > 0) {
    if (c.t1 == c.ce) {
      if (active)
#pragma unroll
        for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = endv_prv_self[lane + z * U];
      return;
    }
    double used[Z], want[Z];
    const double *pe = A.endv + prv + (size_t)(warp + 1) * K;   // next chunk's beta~ at its first SNP
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      used[z] = active ? startv[lane + z * U] : 0.0;
      want[z] = active ? pe[lane + z * U] : 0.0;
    }
    if (!vec_mismatch<Z>(used, want, A.rtol, A.atol)) {
      if (active)
#pragma unroll
        for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = endv_prv_self[lane + z * U];
      return;
    }
#pragma unroll
    for (int z = 0; z < Z; ++z) b[z] = want[z];
    tn = c.t1;
    from_boundary = true;
  } else {
    tn = min(c.t1 + A.warm, c.ce);
    from_boundary = false;
#pragma unroll
    for (int z = 0; z < Z; ++z) b[z] = active ? 1.0 : 0.0;
  }
  if (lane == 0) atomicAdd(A.rerun + A.round, 1u);

  LaneModel<Z> L;
  if (!FROM_PY) load_lane_model<Z>(L, A.M, U, lane, A.het);
  const bool het = active && A.het[lane];

  // stream position q <-> SNP index t = (tn - 1) - q + 1 = tn - q : q = 0 is SNP tn (clamped when tn == ce)
  RecStream<-1> rs;
  rs.init(A.snp, A.txn, smem_all + (threadIdx.x >> 5) * kWarpSmem, tn, c.cs, c.ce - 1, lane);
  SnpRec rec;          // record of SNP t+1 (emission input of the beta step)
  TxnRec tr;           // transition between t and t+1
  double B[Z];         // B_{t+1}
  rs.get(0, rec, tr);
  if (tn < c.ce) {
    if (FROM_PY) emission_from_py<Z>(A.py + (size_t)tn * K, U, lane, active, B);
    else emission<Z>(L, rec, B);
  } else {
#pragma unroll
    for (int z = 0; z < Z; ++z) B[z] = 0.0;
  }

  double st[6][Z];
#pragma unroll
  for (int qq = 0; qq < 6; ++qq)
#pragma unroll
    for (int z = 0; z < Z; ++z) st[qq][z] = 0.0;

  int q = 1;           // step q produces beta~ at SNP t = tn - q
  // ---- warm-up steps: t = tn-1 .. t1 ----
  for (; tn - q >= c.t1; ++q) {
    rs.advance(q);
    SnpRec recn;
    TxnRec trn;
    rs.get(q, recn, trn);                 // record of SNP t (next iteration's emission input)
    double Bn[Z];
    if (FROM_PY) emission_from_py<Z>(A.py + (size_t)(tn - q) * K, U, lane, active, Bn);
    else emission<Z>(L, recn, Bn);
    if (tn - q + 1 < c.ce) bwd_step<Z>(b, B, tr, het, active);
    tr = trn;
#pragma unroll
    for (int z = 0; z < Z; ++z) B[z] = Bn[z];
  }
  // ---- boundary: b[] is beta~_{t1}; normalise what the warm-up produced and publish it ----
  if (c.t1 < c.ce) {
    if (!from_boundary) {
      double loc = 0.0;
#pragma unroll
      for (int z = 0; z < Z; ++z) loc += b[z];
      const double inv = 1.0 / warp_sum(loc);
#pragma unroll
      for (int z = 0; z < Z; ++z) b[z] *= inv;
    }
    if (active)
#pragma unroll
      for (int z = 0; z < Z; ++z) startv[lane + z * U] = b[z];
  }
  // ---- owned steps: t = t1-1 .. t0 ----
  for (; tn - q >= c.t0; ++q) {
    const int t = tn - q;
    rs.advance(q);
    SnpRec recn;
    TxnRec trn;
    rs.get(q, recn, trn);                 // record of SNP t: statistics now, emission input next iteration
    double Bn[Z];
    if (FROM_PY) emission_from_py<Z>(A.py + (size_t)t * K, U, lane, active, Bn);
    else emission<Z>(L, recn, Bn);
    double al[Z];
#pragma unroll
    for (int z = 0; z < Z; ++z) al[z] = active ? A.alpha[(size_t)t * K + lane + z * U] : 0.0;
    if (t + 1 < c.ce) bwd_step<Z>(b, B, tr, het, active);
    // ---- posterior gamma_t ~ alpha~_t * beta~_t ----
    double p[Z];
    double loc = 0.0;
#pragma unroll
    for (int z = 0; z < Z; ++z) { p[z] = al[z] * b[z]; loc += p[z]; }
    const double inv = 1.0 / warp_sum(loc);
    double gsum = 0.0;
#pragma unroll
    for (int z = 0; z < Z; ++z) { p[z] *= inv; gsum += p[z]; }
    if (!FROM_PY) {
      const double l2 = recn.lr * recn.lr;
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        st[0][z] += recn.x * p[z];
        st[1][z] += recn.nx * p[z];
        st[2][z] += recn.lr * p[z];
        st[3][z] += recn.lb * p[z];
        st[4][z] += p[z];
        st[5][z] += l2 * p[z];
      }
    } else if (A.loggamma != nullptr && active) {
#pragma unroll
      for (int z = 0; z < Z; ++z) A.loggamma[(size_t)t * K + lane + z * U] = log(p[z]);
    }
    const bool first = (t == c.cs);
    if (WRITE_POST || first) {
      double rz[Z];
#pragma unroll
      for (int z = 0; z < Z; ++z) rz[z] = warp_sum(p[z]);
      if (WRITE_POST && A.rhoG != nullptr) {
        if (active) A.rhoG[(size_t)t * U + lane] = gsum;
#pragma unroll
        for (int z = 0; z < Z; ++z)
          if (lane == z) A.rhoZ[(size_t)t * Z + z] = rz[z];
      }
      if (first && A.rhoG0 != nullptr) {
        if (active) A.rhoG0[(size_t)c.chr * U + lane] = gsum;
#pragma unroll
        for (int z = 0; z < Z; ++z)
          if (lane == z) A.rhoZ0[(size_t)c.chr * Z + z] = rz[z];
      }
    }
    tr = trn;
#pragma unroll
    for (int z = 0; z < Z; ++z) B[z] = Bn[z];
  }
  // boundary vector for the previous chunk: beta~ at this chunk's first SNP, normalised
  {
    double loc = 0.0;
#pragma unroll
    for (int z = 0; z < Z; ++z) loc += b[z];
    const double inv = 1.0 / warp_sum(loc);
    if (active)
#pragma unroll
      for (int z = 0; z < Z; ++z) endv_cur[lane + z * U] = b[z] * inv;
  }
  if (!FROM_PY && active) {
    double *dst = A.part + (size_t)warp * 6 * K;
#pragma unroll
    for (int qq = 0; qq < 6; ++qq)
#pragma unroll
      for (int z = 0; z < Z; ++z) dst[qq * K + lane + z * U] = st[qq][z];
  }
}

// ------------------------------------------------------------------------------------------
// Deterministic reduction of the per-chunk partial sums and log-likelihood pieces.
// Fixed two-level order (slices of chunks, then slices in order) => bitwise reproducible.
// out layout: [6*K stats | nChr loglik]
// ------------------------------------------------------------------------------------------
__global__ void reduce_slices_kernel(const double *__restrict__ part, int nChunks, int width, int slice,
                                     double *__restrict__ tmp) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (k >= width) return;
  const int c0 = s * slice, c1 = min(c0 + slice, nChunks);
  double acc = 0.0;
  for (int c = c0; c < c1; ++c) acc += part[(size_t)c * width + k];
  tmp[(size_t)s * width + k] = acc;
}

__global__ void reduce_final_kernel(const double *__restrict__ tmp, int nSlices, int width,
                                    const double *__restrict__ chunk_ll, const Chunk *__restrict__ chunks,
                                    int nChunks, int nChr, double *__restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < width) {
    double acc = 0.0;
    for (int s = 0; s < nSlices; ++s) acc += tmp[(size_t)s * width + k];
    out[k] = acc;
  } else if (k < width + nChr) {
    const int chr = k - width;
    double acc = 0.0;
    for (int c = 0; c < nChunks; ++c)
      if (chunks[c].chr == chr) acc += chunk_ll[c];
    out[k] = acc;
  }
}

// structured transition coefficients from the host-exact rhoG/rhoZ (parameter independent)
__global__ void txn_coeff_kernel(const double *__restrict__ rhoG, const double *__restrict__ rhoZ, int N, int K, int U,
                                 int Z, int h, TxnRec *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const double rG = rhoG[t], rZ = rhoZ[t];
  const double oG = (1.0 - rG) / ((double)K - 1.0);
  const double nz = 1.0 - rZ;
  const double q = rZ + (double)(Z - 1) * nz;
  const double rs_non = rG * q + oG * ((double)(h * Z) * rZ + (double)(U - 1 - h) * q);
  const double rs_het = rG * (double)Z * rZ + oG * ((double)((h - 1) * Z) * rZ + (double)(U - h) * q);
  TxnRec r;
  r.k0 = oG * nz;
  r.k1 = oG * (rZ - nz);
  r.k2 = (rG - oG) * nz;
  r.k3 = (rG - oG) * (rZ - nz);
  r.cB = oG * rZ;
  r.k4 = (rG - oG) * rZ;
  r.irn = 1.0 / rs_non;
  r.irh = (h > 0) ? 1.0 / rs_het : 0.0;
  out[t] = r;
}

// ------------------------------------------------------------------------------------------
// Viterbi.  One CTA per chromosome, one thread per destination state; delta lives in shared
// memory, the per-step log-transition values come from a host-built table that reproduces the
// reference's arithmetic bit for bit (sequential row sums, host libm log).  Every addition is a
// single rounded __dadd_rn (no FMA), ties go to the lowest source index.
// reference: src/viterbiC_clonalCN.c:90-128, :210-229
// ------------------------------------------------------------------------------------------
struct VitArgs {
  const SnpRec *snp;
  const int *chr_start;        // [nChr+1]
  const int *txn_id;           // [N] index of the distinct (rhoG,rhoZ) pair between t-1 and t
  const double *ltab;          // [nDistinct][K][4] log(raw_class / rowsum_i), class = 2*sameG + sameZ
  ModelConsts M;
  const unsigned char *het;    // [U]
  const double *py;            // legacy: [K x N] log emissions, else nullptr
  int U, Z, K;
  unsigned char *psi;          // [N][K] back-pointers (K <= 256)
  int *path;                   // [N] 1-based
};

// exact log emission in the reference's association order; no contraction
__device__ __forceinline__ double emission_exact(const SnpRec &r, double lmu, double l1mu, double muC, double cN,
                                                 double tv) {
  const double lik = __dadd_rn(__dmul_rn(r.x, lmu), __dmul_rn(r.nx, l1mu));
  const double pyR = __dadd_rn(r.lb, lik);
  const double d = __dsub_rn(r.lr, muC);
  const double pyC = __dadd_rn(cN, __ddiv_rn(-__dmul_rn(d, d), tv));
  return __dadd_rn(__dadd_rn(pyR, pyC), 1e-300);
}

template <bool FROM_PY>
__global__ void viterbi_forward_kernel(VitArgs A) {
  extern __shared__ double sm[];
  const int K = A.K, U = A.U;
  double *delta0 = sm, *delta1 = sm + K;
  double *lt = sm + 2 * K;            // [K][4] current transition table
  __shared__ int cur_id;
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  const int j = threadIdx.x;
  const bool on = j < K;
  const int gj = on ? j % U : 0, zj = on ? j / U : 0;
  const bool hetj = on && A.het[gj];
  double lmu = 0, l1mu = 0, muC = 0, cN = 0, tv = 1;
  if (on && !FROM_PY) {
    lmu = A.M.lmu[j]; l1mu = A.M.l1mu[j]; muC = A.M.muC[j]; cN = A.M.cN[gj]; tv = A.M.tv[gj];
  }
  if (threadIdx.x == 0) cur_id = -1;
  if (cs >= ce) return;
  {
    double e0 = 0.0;
    if (on) e0 = FROM_PY ? A.py[(size_t)cs * K + j] : emission_exact(A.snp[cs], lmu, l1mu, muC, cN, tv);
    if (on) { delta0[j] = __dadd_rn(A.M.logpi[j], e0); A.psi[(size_t)cs * K + j] = 0; }
  }
  __syncthreads();
  double *dp = delta0, *dc = delta1;
  for (int t = cs + 1; t < ce; ++t) {
    const int id = A.txn_id[t];
    if (id != cur_id) {            // uniform across the CTA (cur_id read after the barrier below)
      for (int q = threadIdx.x; q < 4 * K; q += blockDim.x) lt[q] = A.ltab[(size_t)id * 4 * K + q];
    }
    double e = 0.0;
    if (on) e = FROM_PY ? A.py[(size_t)t * K + j] : emission_exact(A.snp[t], lmu, l1mu, muC, cN, tv);
    __syncthreads();
    if (threadIdx.x == 0) cur_id = id;
    if (on) {
      double best = 0.0;
      int arg = 0;
      int gi = 0, zi = 0;
      for (int i = 0; i < K; ++i) {
        const int cls = ((gi == gj) ? 2 : 0) | ((zi == zj || hetj) ? 1 : 0);
        const double cand = __dadd_rn(dp[i], lt[i * 4 + cls]);
        if (i == 0 || best < cand) { best = cand; arg = i; }
        if (++gi == U) { gi = 0; ++zi; }
      }
      dc[j] = __dadd_rn(best, e);
      A.psi[(size_t)t * K + j] = (unsigned char)arg;
    }
    __syncthreads();
    double *sw = dp; dp = dc; dc = sw;
  }
  // first-index arg-max of the last delta (viterbiC_clonalCN.c:118) and its publication
  if (threadIdx.x == 0) {
    int arg = 0;
    for (int i = 1; i < K; ++i)
      if (dp[arg] < dp[i]) arg = i;
    A.path[ce - 1] = arg;            // 0-based for now; the backtrace adds 1
  }
}

// back-trace: one CTA per chromosome stages psi rows through shared memory in coalesced tiles,
// thread 0 chases the pointers inside the tile (a dependent chain of shared-memory loads
// instead of a dependent chain of HBM loads).
__global__ void viterbi_backtrace_kernel(VitArgs A) {
  extern __shared__ unsigned char tile[];
  __shared__ int s_state;
  const int TILE = 128;
  const int K = A.K;
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  if (cs >= ce) return;
  if (threadIdx.x == 0) {
    s_state = A.path[ce - 1];
    A.path[ce - 1] = s_state + 1;
  }
  __syncthreads();
  // walk t = ce-1 .. cs+1: path[t-1] = psi[t][path[t]]
  for (int hi = ce - 1; hi > cs; hi -= TILE) {
    const int lo = max(hi - TILE + 1, cs + 1);          // rows lo..hi
    const size_t base = (size_t)lo * K, n = (size_t)(hi - lo + 1) * K;
    for (size_t q = threadIdx.x; q < n; q += blockDim.x) tile[q] = A.psi[base + q];
    __syncthreads();
    if (threadIdx.x == 0) {
      int state = s_state;
      for (int t = hi; t >= lo; --t) {
        state = tile[(size_t)(t - lo) * K + state];
        A.path[t - 1] = state + 1;
      }
      s_state = state;
    }
    __syncthreads();
  }
}

}  // namespace titan
