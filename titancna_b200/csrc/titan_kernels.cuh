// titan_kernels.cuh -- sm_100a kernels of the TitanCNA hot path other than the forward-backward engine
// (titan_fb2.cu): deterministic reductions, transition coefficients, the dense compatibility kernels and Viterbi.
#pragma once
#include "titan_common.cuh"

namespace titan {

// ------------------------------------------------------------------------------------------
// Deterministic reduction of the per-chunk partial sums and log-likelihood pieces, one launch.
// Blocks [0, ceil(width/32)): 32 columns x 8 chunk slices per block, every thread adds its slice in chunk order,
// thread (col, 0) adds the eight slice sums in order.  Last block: one thread per chromosome adds the chunk
// log-likelihood pieces of its own chunk range.  Fixed order => bitwise reproducible.
// out layout: [6*K stats | nChr loglik]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_kernel(const double *__restrict__ part, int nChunks, int width,
                                                     const double *__restrict__ chunk_ll, const int *__restrict__ chr_chunk0,
                                                     int nChr, double *__restrict__ out) {
  __shared__ double sm[8][33];
  const int nColBlocks = (width + 31) / 32;
  if ((int)blockIdx.x < nColBlocks) {
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + col;
    const int per = (nChunks + 7) / 8;
    const int c0 = sl * per, c1 = min(c0 + per, nChunks);
    double acc = 0.0;
    if (k < width)
      for (int c = c0; c < c1; ++c) acc += part[(size_t)c * width + k];
    sm[sl][col] = acc;
    __syncthreads();
    if (sl == 0 && k < width) {
      double tot = sm[0][col];
#pragma unroll
      for (int q = 1; q < 8; ++q) tot += sm[q][col];
      out[k] = tot;
    }
  } else {
    for (int chr = threadIdx.x; chr < nChr; chr += blockDim.x) {
      double acc = 0.0;
      for (int c = chr_chunk0[chr]; c < chr_chunk0[chr + 1]; ++c) acc += chunk_ll[c];
      out[width + chr] = acc;
    }
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
  const double k0 = oG * nz, k1 = oG * (rZ - nz), k2 = (rG - oG) * nz, k3 = (rG - oG) * (rZ - nz), cB = oG * rZ, k4 = (rG - oG) * rZ;
  TxnRec r;
  r.h[0] = TxnHalf{1.0 / rs_non, k0, k2, k1, k3, k0, cB, k1};
  r.h[1] = TxnHalf{(h > 0) ? 1.0 / rs_het : 0.0, cB, k4, 0.0, 0.0, k0, cB, k1};
  out[t] = r;
}

// ------------------------------------------------------------------------------------------
// Dense forward-backward for the compatibility path: ANY state layout the reference's C accepts
// (outlier state, repeated zygosity keys, K not a multiple of the cluster count), scaled linear space.
// One CTA per call (one chromosome), one thread per state; the transition between two SNPs is the
// reference's: raw = {oG, rhoG}[same zygosity] * {1-rhoZ, rhoZ}[same cluster or destination HET], rows
// divided by their sums (src/fwd_backC_clonalCN.c:215-301) -- the class of every (i, j) pair comes from
// a host-built K x K table, so no structure is assumed.  O(K^2) per SNP like the reference, but without
// its 2 K^2 log() + 2 K^2 exp() per SNP.  Also serves as the cancellation-free cross-check of the
// structured kernels (tests).
// ------------------------------------------------------------------------------------------
struct DenseArgs {
  const double *py;            // [T][K] log emissions
  const double *logpi;         // [K]
  const double *rhoG, *rhoZ;   // [T] host-exact, entry t = between t-1 and t
  const unsigned char *cls;    // [K][K] class of i -> j: 2*sameZygosity + (sameCluster || destHET)
  int K, T;
  double *alpha;               // [T][K] scratch
  double *loggamma;            // [T][K] out
  double *loglik;              // [1] out
};

__device__ __forceinline__ double block_sum(double v, double *red, int nwarps) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < nwarps; ++w) tot += red[w];
  __syncthreads();
  return tot;
}

__device__ __forceinline__ double block_max(double v, double *red, int nwarps) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = red[0];
  for (int w = 1; w < nwarps; ++w) m = fmax(m, red[w]);
  __syncthreads();
  return m;
}

__global__ void __launch_bounds__(1024) dense_fwd_back_kernel(DenseArgs A) {
  extern __shared__ double dsm[];
  const int K = A.K, T = A.T, j = threadIdx.x, nw = (blockDim.x + 31) >> 5;
  double *w = dsm;                 // [K] alpha / rowsum (forward) or beta * B (backward)
  double *red = dsm + K;           // [32]
  const bool on = j < K;
  const unsigned char *crow = A.cls + (size_t)(on ? j : 0) * K;      // classes of transitions j -> *
  double ll = 0.0;
  // ---- forward ----
  double a;
  {
    const double p0 = on ? A.py[j] : -CUDART_INF;
    const double lp = on ? A.logpi[j] : -CUDART_INF;
    const double mx = block_max(p0 + lp, red, nw);
    a = on ? exp(p0 + lp - mx) : 0.0;
    const double tot = block_sum(a, red, nw);
    a /= tot;
    ll = mx + log(tot);
    if (on) A.alpha[j] = a;
  }
  for (int t = 1; t < T; ++t) {
    const double rG = A.rhoG[t], rZ = A.rhoZ[t];
    const double oG = (1.0 - rG) / ((double)K - 1.0);
    const double val[4] = {oG * (1.0 - rZ), oG * rZ, rG * (1.0 - rZ), rG * rZ};
    double rs = 0.0;
    if (on)
      for (int i = 0; i < K; ++i) rs += val[crow[i]];
    if (on) w[j] = rs > 0 ? a / rs : a;
    const double pyj = on ? A.py[(size_t)t * K + j] : -CUDART_INF;
    const double mx = block_max(pyj, red, nw);       // (also orders the w[] writes before the reads below)
    double m = 0.0;
    if (on)
      for (int i = 0; i < K; ++i) m += w[i] * val[A.cls[(size_t)i * K + j]];
    a = on ? m * exp(pyj - mx) : 0.0;
    const double tot = block_sum(a, red, nw);
    a /= tot;
    ll += mx + log(tot);
    if (on) A.alpha[(size_t)t * K + j] = a;
  }
  if (j == 0) A.loglik[0] = ll;
  // ---- backward + posterior ----
  double b = on ? 1.0 : 0.0;
  if (on) A.loggamma[(size_t)(T - 1) * K + j] = log(a);
  for (int t = T - 2; t >= 0; --t) {
    const double rG = A.rhoG[t + 1], rZ = A.rhoZ[t + 1];
    const double oG = (1.0 - rG) / ((double)K - 1.0);
    const double val[4] = {oG * (1.0 - rZ), oG * rZ, rG * (1.0 - rZ), rG * rZ};
    const double pyj = on ? A.py[(size_t)(t + 1) * K + j] : -CUDART_INF;
    const double mx = block_max(pyj, red, nw);
    if (on) w[j] = b * exp(pyj - mx);
    __syncthreads();
    double rs = 0.0, o = 0.0;
    if (on)
      for (int i = 0; i < K; ++i) { const double v = val[crow[i]]; rs += v; o += v * w[i]; }
    b = on ? (rs > 0 ? o / rs : o) : 0.0;
    const double tot = block_sum(b, red, nw);
    b /= tot;
    const double al = on ? A.alpha[(size_t)t * K + j] : 0.0;
    const double g = al * b;
    const double gt = block_sum(g, red, nw);
    if (on) A.loggamma[(size_t)t * K + j] = log(g / gt);
  }
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
  const unsigned char *cls;    // optional [K][K] transition classes (generic legacy path: outlier state, repeated zygosity keys)
  int U, Z, K;
  unsigned char *psi;          // [N][K] back-pointers (K <= 256)
  int *path;                   // [N] 1-based
  // tiled back-trace
  int nChr, bt_rows;           // rows of psi per tile
  const int *tile_first;       // [nChr+1] first tile of every chromosome
  unsigned char *comp;         // [nTiles][K] state entering a tile -> state leaving it
  unsigned char *entry;        // [nTiles] state at the top row of a tile
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

// Gaussian allele model: bivariateNormalpdflog (R/hmmClonal.R:550-559) in R's association order, single roundings;
// the Viterbi wrapper adds no pseudoCounts to it (:444-447).  r.x = ref / tumDepth (host division).
__device__ __forceinline__ double emission_exact_gauss(const SnpRec &r, double muR, double muC, double negc, double varR,
                                                       double varC, double sq, double l0, double rho2) {
  const double d1 = __dsub_rn(r.x, muR), d2 = __dsub_rn(r.lr, muC);
  const double l1 = __ddiv_rn(__dmul_rn(d1, d1), varR);
  const double l2 = __ddiv_rn(__dmul_rn(d2, d2), varC);
  const double l3 = __ddiv_rn(__dmul_rn(__dmul_rn(rho2, d1), d2), sq);
  return __dsub_rn(negc, __dmul_rn(l0, __dsub_rn(__dadd_rn(l1, l2), l3)));
}

// Thread layout: P = blockDim/K' threads per destination state (P a power of two <= 32, consecutive
// lanes), each scanning a contiguous ascending slice of the sources.  Per step:
//   phase A  thread i < K forms the four class values v[c][i] = delta[i] (+) logA_class_c(i)  -- 4K
//            roundings instead of K^2: every destination sharing class c with source i would add
//            the very same two doubles;
//   phase B  destination j takes the first-index arg-max over i of v[class(i,j)][i]: slice-local
//            strict '<' scans, then a butterfly over the P slices ordered by (value desc, index asc),
//            which is exactly the reference's left-to-right scan with strict '<' (:210-229).
template <bool FROM_PY>
__global__ void __launch_bounds__(1024) viterbi_forward_kernel(VitArgs A, int P) {
  extern __shared__ double sm[];
  const int K = A.K, U = A.U, K4 = 4 * A.K;
  double *delta = sm;                 // [K]
  double *v = sm + K;                 // [4][K] class values of the current step
  double *ltb = sm + 5 * K;           // [2][K][4] transition tables of the current and the next step
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  const int tid = threadIdx.x;
  const int j = tid / P, p = tid % P;               // destination state, slice
  const bool on = j < K;
  const int gj = on ? j % U : 0, zj = on ? j / U : 0;
  const bool hetj = on && A.het[gj];
  const bool lead = on && p == 0;
  const int per = (K + P - 1) / P;
  const int i0 = min(p * per, K), i1 = min(i0 + per, K);
  const int nthr = blockDim.x;
  const int cnt = on ? max(i1 - i0, 0) : 0;
  unsigned short *offs = reinterpret_cast<unsigned short *>(sm + 13 * K);   // [per][blockDim]
  {
    int gi = i0 % U, zi = i0 / U;
    for (int n = 0; n < cnt; ++n) {
      const int cls = A.cls ? A.cls[(size_t)(i0 + n) * K + j] : (((gi == gj) ? 2 : 0) | ((zi == zj || hetj) ? 1 : 0));
      offs[n * nthr + tid] = (unsigned short)(cls * K + i0 + n);
      if (++gi == U) { gi = 0; ++zi; }
    }
  }
  double lmu = 0, l1mu = 0, muC = 0, cN = 0, tv = 1;
  if (lead && !FROM_PY) {
    lmu = A.M.lmu[j]; l1mu = A.M.l1mu[j]; muC = A.M.muC[j]; cN = A.M.cN[gj]; tv = A.M.tv[gj];
  }
  if (cs >= ce) return;
  if (lead) {
    const double e0 = FROM_PY ? A.py[(size_t)cs * K + j] : emission_exact(A.snp[cs], lmu, l1mu, muC, cN, tv);
    delta[j] = __dadd_rn(A.M.logpi[j], e0);
    A.psi[(size_t)cs * K + j] = 0;
  }
  // Nothing of step t+1 except delta depends on the recursion: its transition table (adjacent SNP
  // gaps differ, so the table changes at almost every step), its table id and its emission are
  // fetched one step ahead and their global-memory latency hides behind phases A and B.
  if (cs + 1 < ce) {
    const int id0 = A.txn_id[cs + 1];
    for (int q = tid; q < K4; q += blockDim.x) ltb[q] = A.ltab[(size_t)id0 * K4 + q];
  }
  int id_n = (cs + 2 < ce) ? A.txn_id[cs + 2] : 0;
  double e = 0.0;
  if (lead && cs + 1 < ce)
    e = FROM_PY ? A.py[(size_t)(cs + 1) * K + j] : emission_exact(A.snp[cs + 1], lmu, l1mu, muC, cN, tv);
  __syncthreads();
  for (int t = cs + 1; t < ce; ++t) {
    const int cur = (t - cs - 1) & 1;
    const double *lt = ltb + cur * K4;
    const int tn = min(t + 1, ce - 1), tnn = min(t + 2, ce - 1);
    // prefetch for step t+1 (registers now, shared memory after phase B)
    double ltn[2] = {0.0, 0.0};
    if (tid < K4) ltn[0] = A.ltab[(size_t)id_n * K4 + tid];
    if (tid + (int)blockDim.x < K4) ltn[1] = A.ltab[(size_t)id_n * K4 + tid + blockDim.x];
    const int id_nn = A.txn_id[tnn];
    SnpRec rec_n;
    double py_n = 0.0;
    if (lead) {
      if (FROM_PY) py_n = A.py[(size_t)tn * K + j];
      else rec_n = A.snp[tn];
    }
    // phase A
    if (tid < K) {
      const double d = delta[tid];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c * K + tid] = __dadd_rn(d, lt[tid * 4 + c]);
    }
    __syncthreads();
    // phase B
    double best = -CUDART_INF;
    int arg = 0x7fffffff;
    if (cnt > 0) {
      // offs[n] = class(i0+n, j)*K + (i0+n): step-invariant, tabulated once before the loop
      best = v[offs[tid]];
      arg = i0;
      for (int n = 1; n < cnt; ++n) {
        const double cand = v[offs[n * nthr + tid]];
        if (best < cand) { best = cand; arg = i0 + n; }
      }
    }
    for (int o = 1; o < P; o <<= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    // delta was last read in phase A (before the barrier above), so it can be overwritten here
    if (lead) {
      delta[j] = __dadd_rn(best, e);
      A.psi[(size_t)t * K + j] = (unsigned char)arg;
      e = FROM_PY ? py_n : emission_exact(rec_n, lmu, l1mu, muC, cN, tv);
    }
    double *ltw = ltb + (cur ^ 1) * K4;
    if (tid < K4) ltw[tid] = ltn[0];
    if (tid + (int)blockDim.x < K4) ltw[tid + blockDim.x] = ltn[1];
    id_n = id_nn;
    __syncthreads();
  }
  // first-index arg-max of the last delta (viterbiC_clonalCN.c:118) and its publication
  if (tid == 0) {
    int arg = 0;
    for (int i = 1; i < K; ++i)
      if (delta[arg] < delta[i]) arg = i;
    A.path[ce - 1] = arg;            // 0-based for now; the backtrace adds 1
  }
}

// ------------------------------------------------------------------------------------------
// Exact log emissions of every (SNP, state), materialised once per decode by a fully parallel
// kernel so that the sequential Viterbi kernel below carries no emission arithmetic at all.
// reference: R/hmmClonal.R:449-459, :617-634 (association order of SURVEY.md A.3)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) viterbi_emission_kernel(const SnpRec *__restrict__ snp, ModelConsts M, int U, int K,
                                                               long long total, double *__restrict__ py) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
    const long long t = q / K;
    const int j = (int)(q - t * K), g = j % U;
    const SnpRec r = snp[t];
    py[q] = M.gauss ? emission_exact_gauss(r, M.lmu[j], M.muC[j], M.cN[g], M.vR[g], M.tv[g], M.sq[g], M.l0, M.rho2)
                    : emission_exact(r, M.lmu[j], M.l1mu[j], M.muC[j], M.cN[g], M.tv[g]);
  }
}

// first-index arg-max over compile-time positions [LO, HI) of a register array: the right half wins
// only on a strictly larger value, which is the reference's left-to-right scan with strict '<'
template <int LO, int HI>
struct ArgmaxTree {
  template <int PER>
  static __device__ __forceinline__ void run(const double (&c)[PER], double &bv, int &bi) {
    constexpr int MID = (LO + HI) / 2;
    double lv, rv;
    int li, ri;
    ArgmaxTree<LO, MID>::run(c, lv, li);
    ArgmaxTree<MID, HI>::run(c, rv, ri);
    const bool r = lv < rv;
    bv = r ? rv : lv;
    bi = r ? ri : li;
  }
};
template <int LO>
struct ArgmaxTree<LO, LO + 1> {
  template <int PER>
  static __device__ __forceinline__ void run(const double (&c)[PER], double &bv, int &bi) { bv = c[LO]; bi = LO; }
};

// Latency-lean Viterbi step (same arithmetic, same ordering as viterbi_forward_kernel):
//   * the shared-memory addresses of a thread's PER candidate class values live in REGISTERS, so
//     the slice scan is PER independent loads (register + immediate addressing, the two value
//     buffers sit a compile-time stride apart) followed by a compare tree of depth log2(PER),
//     instead of a chain of PER dependent (offset load -> value load -> compare) links;
//   * the lead thread of destination j turns its fresh delta_t[j] straight into the four class
//     values of SOURCE j for step t+1 (double-buffered, [K][4] so they are two 16-byte stores):
//     a step has ONE barrier and delta never makes a round trip through shared memory;
//   * log emissions come from the materialised py matrix and the transition rows of state j from the
//     per-gap tables, both fetched D steps ahead into registers by the lead thread that uses them.
constexpr int kVitMaxK = 128;                       // largest K served by viterbi_fast_kernel
constexpr int kVitVS = 4 * kVitMaxK + 8;            // doubles per class-value buffer (+ the -inf slot), 16-byte multiple

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void lds_v2f64(unsigned addr, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-step inputs of the sequential Viterbi kernels -- the log emission of state j at step s, the four
// class log-transitions of row j for the move s -> s+1, and the id of the table needed DV steps later --
// stream through a 4-slot shared-memory ring filled by cp.async: no register is the target of a global
// load inside the step loop, so no scoreboard wait on an in-flight load ever lands on the critical path
// (with register prefetch the loop back-edge waited for the youngest load: 25 % of the step time).
constexpr int DV = 6;                    // prefetch distance in steps; ring slot of step s is (s - cs - 1) & 7
constexpr int RV = 8;                    // ring slots
struct VitRing {
  double *L;    // [RV][K][4]
  double *e;    // [RV][K]
  int *id;      // [RV][K]   txn_id[s + DV + 1], consumed when group s + DV is issued
  __device__ __forceinline__ void carve(double *base, int K) {
    L = base; e = base + 4 * RV * (size_t)K; id = reinterpret_cast<int *>(base + 5 * RV * (size_t)K);
  }
  static __host__ __device__ constexpr size_t bytes(int K) { return (size_t)K * RV * (4 * 8 + 8 + 4); }
  // one commit group: everything state j needs at step s
  __device__ __forceinline__ void issue(const VitArgs &A, int slot, int s, int last, int j, int row_id) const {
    const int K = A.K;
    const int sc = min(s, last);
    cp_async<8>(e + slot * K + j, A.py + (size_t)sc * K + j);
    const double *row = A.ltab + (size_t)row_id * 4 * K + (size_t)j * 4;
    cp_async<16>(L + ((size_t)slot * K + j) * 4, row);
    cp_async<16>(L + ((size_t)slot * K + j) * 4 + 2, row + 2);
    cp_async<4>(id + slot * K + j, A.txn_id + min(s + DV + 1, last));
    cp_async_commit();
  }
  // groups of the first DV steps (s = cs+1 .. cs+DV), table ids read directly
  __device__ __forceinline__ void prologue(const VitArgs &A, int cs, int last, int j) const {
#pragma unroll
    for (int u = 0; u < DV; ++u) {
      const int s = cs + 1 + u;
      issue(A, u, s, last, j, A.txn_id[min(s + 1, last)]);
    }
    cp_async_wait<DV - 1>();
  }
};


template <int PER, int P, int MAXT>
__global__ void __launch_bounds__(MAXT) viterbi_fast_kernel(VitArgs A) {
  extern __shared__ __align__(16) double sm[];
  constexpr int D = 4;                 // unroll of the step loop = ring slots; even, so the buffer parity of a step is static
  constexpr unsigned VSB = kVitVS * sizeof(double);
  const int K = A.K, U = A.U, K4 = 4 * A.K;
  double *vbuf = sm;                   // [2][kVitVS]: class values [K][4] + a -inf slot at 4K
  double *dfin = sm + 2 * kVitVS;      // [K] last delta
  VitRing ring;
  ring.carve(sm + 2 * kVitVS + kVitMaxK, A.K);
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  if (cs >= ce) return;
  const int tid = threadIdx.x;
  const int j = tid / P, p = tid % P;
  const bool on = j < K;
  const int gj = on ? j % U : 0, zj = on ? j / U : 0;
  const bool hetj = on && A.het[gj];
  const bool lead = on && p == 0;
  const int per = (K + P - 1) / P;
  const int i0 = min(p * per, K), i1 = min(i0 + per, K);
  const int cnt = on ? max(i1 - i0, 0) : 0;
  const unsigned vbase = (unsigned)__cvta_generic_to_shared(vbuf);
  unsigned adr[PER];                   // shared-memory byte address (buffer 0) of candidate n
  {
    int gi = i0 % U, zi = i0 / U;
#pragma unroll
    for (int n = 0; n < PER; ++n) {
      int o = K4;                      // the -inf slot
      if (n < cnt) {
        const int cls = A.cls ? A.cls[(size_t)(i0 + n) * K + j] : (((gi == gj) ? 2 : 0) | ((zi == zj || hetj) ? 1 : 0));
        o = (i0 + n) * 4 + cls;
        if (++gi == U) { gi = 0; ++zi; }
      }
      adr[n] = vbase + 8u * (unsigned)o;
    }
  }
  if (tid < 2) vbuf[tid * kVitVS + K4] = -CUDART_INF;
  const int last = ce - 1;
  const double *pyj = A.py + j;
  const double *ltj = A.ltab + (size_t)j * 4;
  if (lead) ring.prologue(A, cs, last, j);
  // ---- t = cs ----
  if (lead) {
    const double d0 = __dadd_rn(A.M.logpi[j], pyj[(size_t)cs * K]);
    A.psi[(size_t)cs * K + j] = 0;
    dfin[j] = d0;
    if (cs + 1 < ce) {
      const double2 *row = reinterpret_cast<const double2 *>(ltj + (size_t)A.txn_id[cs + 1] * K4);
      const double2 a = row[0], b = row[1];
      double2 *dst = reinterpret_cast<double2 *>(vbuf + 4 * j);
      dst[0] = make_double2(__dadd_rn(d0, a.x), __dadd_rn(d0, a.y));
      dst[1] = make_double2(__dadd_rn(d0, b.x), __dadd_rn(d0, b.y));
    }
  }
  __syncthreads();
  for (int t = cs + 1; t < ce; t += D) {
#pragma unroll
    for (int u = 0; u < D; ++u) {
      const int tt = t + u;
      if (tt >= ce) break;
      constexpr unsigned zero = 0;
      const unsigned rd = (u & 1) ? VSB : zero;       // step t+u reads buffer (u & 1), writes the other one
      // this step's ring slot (filled by this very thread): off the critical path, consumed after the scan
      double e_u = 0.0;
      double2 la_u = make_double2(0, 0), lb_u = make_double2(0, 0);
      int id_u = 0;
      const int slot = (tt - cs - 1) & (RV - 1);
      if (lead) {
        e_u = ring.e[slot * K + j];
        const double2 *rl = reinterpret_cast<const double2 *>(ring.L + ((size_t)slot * K + j) * 4);
        la_u = rl[0]; lb_u = rl[1];
        id_u = ring.id[slot * K + j];
      }
      double c[PER];
#pragma unroll
      for (int n = 0; n < PER; ++n) c[n] = lds_f64(adr[n] + rd);
      double best;
      int pos;
      ArgmaxTree<0, PER>::run(c, best, pos);
      int arg = i0 + pos;
#pragma unroll
      for (int o = 1; o < P; o <<= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        const bool take = ob > best || (ob == best && oa < arg);
        best = take ? ob : best;
        arg = take ? oa : arg;
      }
      if (lead) {
        const double d = __dadd_rn(best, e_u);
        double2 *dst = reinterpret_cast<double2 *>(vbuf + ((u & 1) ? 0 : kVitVS) + 4 * j);
        dst[0] = make_double2(__dadd_rn(d, la_u.x), __dadd_rn(d, la_u.y));
        dst[1] = make_double2(__dadd_rn(d, lb_u.x), __dadd_rn(d, lb_u.y));
        A.psi[(size_t)tt * K + j] = (unsigned char)arg;
        if (tt == last) dfin[j] = d;
        ring.issue(A, (slot + DV) & (RV - 1), tt + DV, last, j, id_u);
        cp_async_wait<DV - 1>();      // the group of step tt+1 has landed (own copies: visible to this thread)
      }
      __syncthreads();
    }
  }
  // first-index arg-max of the last delta (viterbiC_clonalCN.c:118) and its publication
  if (tid == 0) {
    int arg = 0;
    for (int i = 1; i < K; ++i)
      if (dfin[arg] < dfin[i]) arg = i;
    A.path[ce - 1] = arg;            // 0-based for now; the backtrace adds 1
  }
}

// ------------------------------------------------------------------------------------------
// Cluster-sliced Viterbi step: O(K*U) compares per SNP instead of O(K^2), bit-exact, ONE barrier.
//
// Warp (z', h) scans slice h of the sources of clonal cluster z' (lane = genotype g of the
// DESTINATION).  For a destination of genotype g every source (g', z') offers one of two values,
// depending only on whether the destination sits in cluster z' (or is diploid HET) -- class 1/3 -- or
// not -- class 0/2 -- so the warp produces, per lane, two first-index arg-maxima over its slice:
//   X = max over the slice of (g' == g ? v2 : v0)      for destinations (g, z != z')
//   Y = max over the slice of (g' == g ? v3 : v1)      for destination  (g, z') and HET destinations
// The class values [source][4] sit in a warp-private shared-memory table (all H warps of a cluster
// keep identical copies, so only __syncwarp orders their update); lane g loads one 16-byte pair per
// source through a register-resident address that points 16 bytes further for its own genotype.
// Slices are contiguous and ascending in source index, so both the slice scan and the gather over
// the Z*H slice results are compare trees in which the right operand wins only when strictly larger:
// the reference's left-to-right scan with strict '<' (src/viterbiC_clonalCN.c:210-229).
// ------------------------------------------------------------------------------------------
struct __align__(16) VCand { double v; int i; int pad; };

template <int LO, int HI>
struct PairTree {     // first-index arg-max over candidates [LO, HI) held as (value, index) register arrays
  template <int N>
  static __device__ __forceinline__ void run(const double (&v)[N], const int (&ix)[N], double &bv, int &bi) {
    constexpr int MID = (LO + HI) / 2;
    double lv, rv;
    int li, ri;
    PairTree<LO, MID>::run(v, ix, lv, li);
    PairTree<MID, HI>::run(v, ix, rv, ri);
    const bool r = lv < rv;
    bv = r ? rv : lv;
    bi = r ? ri : li;
  }
};
template <int LO>
struct PairTree<LO, LO + 1> {
  template <int N>
  static __device__ __forceinline__ void run(const double (&v)[N], const int (&ix)[N], double &bv, int &bi) { bv = v[LO]; bi = ix[LO]; }
};

template <int Z, int H, int UP>
__global__ void __launch_bounds__(32 * Z * H) viterbi_cluster_kernel(VitArgs A) {
  extern __shared__ __align__(16) unsigned char vc_sm[];
  constexpr int W = Z * H;                  // warps
  constexpr int D = 4;                      // unroll of the step loop = ring slots; even: the exchange parity of a step is static
  constexpr int TABD = 33 * 4;              // doubles per class-value table: [32 sources][4 classes] + a -inf row
  double *tab_all = reinterpret_cast<double *>(vc_sm);                       // [W][TABD]
  VCand *xch = reinterpret_cast<VCand *>(tab_all + W * TABD);                // [2 parity][2 X/Y][W][32]
  double *dfin = reinterpret_cast<double *>(xch + 2 * 2 * W * 32);           // [K]
  VitRing ring;
  ring.carve(dfin + ((A.K + 1) & ~1), A.K);
  const int K = A.K, U = A.U, K4 = 4 * A.K;
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  if (cs >= ce) return;
  const int warp = threadIdx.x >> 5, g = threadIdx.x & 31;
  const int zs = warp / H, h = warp % H;
  const bool on = g < U;
  const int j = on ? g + zs * U : 0;        // the state this lane owns: source (g, zs) of its table, destination (g, zs)
  const bool het = on && A.het[g];
  const bool writer = on && h == 0;            // stores back-pointers and the last delta
  const bool feeder = on && h == H - 1;        // keeps the input ring filled (the other warp of the cluster when H > 1)
  double *tab = tab_all + warp * TABD;
  const int per = (U + H - 1) / H;
  const int g0 = h * per;
  unsigned adr[UP];
  {
    const unsigned tbase = (unsigned)__cvta_generic_to_shared(tab);
#pragma unroll
    for (int n = 0; n < UP; ++n) {
      const int gs = g0 + n;
      const bool valid = n < per && gs < U;
      adr[n] = tbase + (valid ? 32u * (unsigned)gs + (gs == g ? 16u : 0u) : 32u * 32u);
    }
  }
  if (g < 4) tab[32 * 4 + g] = -CUDART_INF;
  // gather addresses: slice result (z', h') that applies to destination cluster zs
  unsigned gadr[W];
  {
    const unsigned xbase = (unsigned)__cvta_generic_to_shared(xch);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int sel = ((w / H) == zs || het) ? 1 : 0;
      gadr[w] = xbase + 16u * (unsigned)((sel * W + w) * 32 + g);
    }
  }
  constexpr unsigned XPAR = 16u * 2 * W * 32;     // bytes between the two exchange parities
  VCand *xmine = xch + warp * 32 + g;             // parity 0, X; Y sits W*32 entries further
  const int last = ce - 1;
  const double *pyj = A.py + j;
  const double *ltj = A.ltab + (size_t)j * 4;
  if (feeder) ring.prologue(A, cs, last, j);      // the H warps of a cluster share the ring entries of their states
  // ---- t = cs ----
  {
    double d0 = -CUDART_INF;
    double2 a = make_double2(0, 0), b = make_double2(0, 0);
    if (on) {
      d0 = __dadd_rn(A.M.logpi[j], pyj[(size_t)cs * K]);
      if (cs + 1 < ce) {
        const double2 *row = reinterpret_cast<const double2 *>(ltj + (size_t)A.txn_id[cs + 1] * K4);
        a = row[0]; b = row[1];
      }
    }
    if (writer) { A.psi[(size_t)cs * K + j] = 0; dfin[j] = d0; }
    double2 *dst = reinterpret_cast<double2 *>(tab + 4 * g);
    dst[0] = make_double2(__dadd_rn(d0, a.x), __dadd_rn(d0, a.y));
    dst[1] = make_double2(__dadd_rn(d0, b.x), __dadd_rn(d0, b.y));
  }
  __syncthreads();
  for (int t = cs + 1; t < ce; t += D) {
#pragma unroll
    for (int u = 0; u < D; ++u) {
      const int tt = t + u;
      if (tt >= ce) break;
      // ---- slice scan ----
      double c0[UP], c1[UP];
#pragma unroll
      for (int n = 0; n < UP; ++n) lds_v2f64(adr[n], c0[n], c1[n]);
      double xv, yv;
      int xp, yp;
      ArgmaxTree<0, UP>::run(c0, xv, xp);
      ArgmaxTree<0, UP>::run(c1, yv, yp);
      const int ibase = zs * U + g0;
      VCand *xo = xmine + ((u & 1) ? 2 * W * 32 : 0);
      VCand cx, cy;
      cx.v = xv; cx.i = ibase + xp; cx.pad = 0;
      cy.v = yv; cy.i = ibase + yp; cy.pad = 0;
      xo[0] = cx;
      xo[W * 32] = cy;
      __syncthreads();
      // ---- gather over the Z*H slices, in source order ----
      double gv[W];
      int gi[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        double ip;
        lds_v2f64(gadr[w] + ((u & 1) ? XPAR : 0u), gv[w], ip);
        gi[w] = __double2loint(ip);
      }
      // this step's ring slot: landed before the barrier above (writer's wait_group in the previous step)
      const int slot = (tt - cs - 1) & (RV - 1);
      const double e_u = ring.e[slot * K + j];
      const double2 *rl = reinterpret_cast<const double2 *>(ring.L + ((size_t)slot * K + j) * 4);
      const double2 la_u = rl[0], lb_u = rl[1];
      double best;
      int arg;
      PairTree<0, W>::run(gv, gi, best, arg);
      const double d = __dadd_rn(best, e_u);
      double2 *dst = reinterpret_cast<double2 *>(tab + 4 * g);
      dst[0] = make_double2(__dadd_rn(d, la_u.x), __dadd_rn(d, la_u.y));
      dst[1] = make_double2(__dadd_rn(d, lb_u.x), __dadd_rn(d, lb_u.y));
      if (writer) {
        A.psi[(size_t)tt * K + j] = (unsigned char)arg;
        if (tt == last) dfin[j] = d;
      }
      if (feeder) {
        ring.issue(A, (slot + DV) & (RV - 1), tt + DV, last, j, ring.id[slot * K + j]);
        cp_async_wait<DV - 1>();      // the group of step tt+1 has landed; the next barrier publishes it
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // first-index arg-max of the last delta (viterbiC_clonalCN.c:118)
    int arg = 0;
    for (int i = 1; i < K; ++i)
      if (dfin[arg] < dfin[i]) arg = i;
    A.path[ce - 1] = arg;            // 0-based for now; the back-trace adds 1
  }
}

// ------------------------------------------------------------------------------------------
// Warp-reduction Viterbi step (default for U <= 64): the slice scan of viterbi_cluster_kernel without the scan.
//
// Warp z' holds cluster z' with lane = SOURCE genotype g': every lane keeps the four class values of its own state
// in registers.  What a destination of genotype g can take from warp z' is
//   X_g = max over g' of (g' == g ? v2[g'] : v0[g'])      destination in another cluster
//   Y_g = max over g' of (g' == g ? v3[g'] : v1[g'])      destination in cluster z' / diploid-HET destination
// and because the same-genotype class never loses to the other-genotype class of the same source (rhoG >= oG, so
// v2 >= v0 and v3 >= v1 in every row of the host-built table; checked on the host when the table is built), the
// own-genotype substitution cannot evict a maximum it does not replace: X_g is the better of v2[g] (index g) and the
// warp-wide first-index maximum (M0, a0) of v0 -- for g == a0 the rule still holds, v2[a0] >= M0.  The warp-wide
// maximum of an fp64 value is two REDUX.MAX on an order-preserving 64-bit integer key (high word, then low word among
// the lanes that hold the high maximum), the first index a ballot + find-first-set: exact, ties to the lowest source.
// One CTA barrier per step publishes the Z per-warp (X, Y) rows; a dedicated warp keeps the input ring filled.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void f64_key(double v, unsigned &hi, unsigned &lo) {
  const int h = __double2hiint(v);
  const unsigned m = (unsigned)(h >> 31);            // all ones for negative values
  hi = (unsigned)h ^ (m | 0x80000000u);
  lo = (unsigned)__double2loint(v) ^ m;
}
__device__ __forceinline__ double f64_unkey(unsigned hi, unsigned lo) {
  const unsigned m = ~(unsigned)((int)hi >> 31);     // key's top bit set <=> value was non-negative
  return __hiloint2double((int)(hi ^ (m | 0x80000000u)), (int)(lo ^ m));
}
// first-index arg-max over the warp of the values whose keys are (hi, lo)
__device__ __forceinline__ void warp_argmax_key(unsigned hi, unsigned lo, double &mv, int &arg) {
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const bool p = hi == mh;
  const unsigned ml = __reduce_max_sync(0xffffffffu, p ? lo : 0u);
  const unsigned who = __ballot_sync(0xffffffffu, p && lo == ml);
  arg = __ffs((int)who) - 1;
  mv = f64_unkey(mh, ml);
}

// GS = 2 serves tables of 33..64 genotypes: lane l owns genotypes l and l + 32 of its cluster.  Source indices order as
// (slot, lane), so a tie between candidates goes to slot 0 first, then to the lowest lane.
template <int Z, int GS>
__global__ void __launch_bounds__(32 * (Z + 1)) viterbi_redux_kernel(VitArgs A) {
  extern __shared__ __align__(16) unsigned char vr_sm[];
  constexpr int D = 4;                      // unroll of the step loop; even: the exchange parity of a step is static
  constexpr int ROW = 32 * GS;              // entries per exchange row
  VCand *xch = reinterpret_cast<VCand *>(vr_sm);                             // [2 parity][2 X/Y][Z][ROW]
  double *dfin = reinterpret_cast<double *>(xch + 2 * 2 * Z * ROW);          // [K]
  VitRing ring;
  ring.carve(dfin + ((A.K + 1) & ~1), A.K);
  const int K = A.K, U = A.U, K4 = 4 * A.K;
  const int chr = blockIdx.x;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  if (cs >= ce) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int last = ce - 1;
  if (warp == Z) {
    // ---- feeder warp: keeps the ring DV steps ahead for every state (lane l serves states l, l+32, ...) ----
    for (int u = 0; u < DV; ++u) {
      const int s = cs + 1 + u;
      const int rid = A.txn_id[min(s + 1, last)];
      for (int j = lane; j < K; j += 32) {
        const int sc = min(s, last);
        cp_async<8>(ring.e + u * K + j, A.py + (size_t)sc * K + j);
        const double *row = A.ltab + (size_t)rid * 4 * K + (size_t)j * 4;
        cp_async<16>(ring.L + ((size_t)u * K + j) * 4, row);
        cp_async<16>(ring.L + ((size_t)u * K + j) * 4 + 2, row + 2);
      }
      cp_async_commit();
    }
    int rid_next = A.txn_id[min(cs + DV + 2, last)];     // table of the group issued next (step cs + DV + 1)
    cp_async_wait<DV - 1>();                             // group of step cs+1 has landed
    __syncthreads();                                     // publishes the t = cs state and group cs+1
    for (int tt = cs + 1; tt < ce; ++tt) {
      // group tt+DV goes into the slot of step tt-2, which every consumer left before the previous barrier
      const int s = tt + DV, slot = (s - cs - 1) & (RV - 1);
      const int rid = rid_next;
      rid_next = A.txn_id[min(s + 2, last)];
      for (int j = lane; j < K; j += 32) {
        const int sc = min(s, last);
        cp_async<8>(ring.e + slot * K + j, A.py + (size_t)sc * K + j);
        const double *row = A.ltab + (size_t)rid * 4 * K + (size_t)j * 4;
        cp_async<16>(ring.L + ((size_t)slot * K + j) * 4, row);
        cp_async<16>(ring.L + ((size_t)slot * K + j) * 4 + 2, row + 2);
      }
      cp_async_commit();
      cp_async_wait<DV - 1>();                           // group of step tt+1 has landed: the next barrier publishes it
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();                                     // the consumers' closing barrier
    return;
  }
  const int zs = warp;
  const int ibase = zs * U;
  bool on[GS], het[GS];
  int g[GS], j[GS];
  unsigned gadr[GS][Z];         // gather addresses: row (z') of X or Y that applies to destination (g, zs)
  {
    const unsigned xbase = (unsigned)__cvta_generic_to_shared(xch);
#pragma unroll
    for (int s = 0; s < GS; ++s) {
      g[s] = lane + 32 * s;
      on[s] = g[s] < U;
      j[s] = on[s] ? g[s] + zs * U : 0;        // this slot's state: source (g, zs) and destination (g, zs)
      het[s] = on[s] && A.het[g[s]];
#pragma unroll
      for (int w = 0; w < Z; ++w) {
        const int sel = (w == zs || het[s]) ? 1 : 0;
        gadr[s][w] = xbase + 16u * (unsigned)((sel * Z + w) * ROW + g[s]);
      }
    }
  }
  constexpr unsigned XPAR = 16u * 2 * Z * ROW;    // bytes between the two exchange parities
  VCand *xmine = xch + zs * ROW + lane;           // parity 0, X, slot 0; slot 1 sits 32 entries further, Y sits Z*ROW further
  double v0[GS], v1[GS], v2[GS], v3[GS];
  // ---- t = cs ----
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    double d0 = -CUDART_INF;
    double2 a = make_double2(0, 0), b = make_double2(0, 0);
    if (on[s]) {
      d0 = __dadd_rn(A.M.logpi[j[s]], A.py[(size_t)cs * K + j[s]]);
      if (cs + 1 < ce) {
        const double2 *row = reinterpret_cast<const double2 *>(A.ltab + (size_t)j[s] * 4 + (size_t)A.txn_id[cs + 1] * K4);
        a = row[0]; b = row[1];
      }
      A.psi[(size_t)cs * K + j[s]] = 0;
      dfin[j[s]] = d0;
    }
    v0[s] = __dadd_rn(d0, a.x); v1[s] = __dadd_rn(d0, a.y); v2[s] = __dadd_rn(d0, b.x); v3[s] = __dadd_rn(d0, b.y);
  }
  __syncthreads();
  for (int t = cs + 1; t < ce; t += D) {
#pragma unroll
    for (int u = 0; u < D; ++u) {
      const int tt = t + u;
      if (tt >= ce) break;
      // ---- warp-wide first-index maxima of the two other-genotype classes ----
      double m0, m1;
      int a0, a1;
      {
        unsigned h0, l0, h1, l1;
        f64_key(v0[0], h0, l0);
        f64_key(v1[0], h1, l1);
        if (!on[0]) { h0 = l0 = h1 = l1 = 0u; }
        int s0 = 0, s1 = 0;                  // slot of this lane's best candidate (slot 0 wins ties: lower index)
        if (GS > 1) {
          unsigned h, l;
          f64_key(v0[GS - 1], h, l);
          if (on[GS - 1] && (h > h0 || (h == h0 && l > l0))) { h0 = h; l0 = l; s0 = 1; }
          f64_key(v1[GS - 1], h, l);
          if (on[GS - 1] && (h > h1 || (h == h1 && l > l1))) { h1 = h; l1 = l; s1 = 1; }
        }
        if (GS == 1) {
          warp_argmax_key(h0, l0, m0, a0);
          warp_argmax_key(h1, l1, m1, a1);
        } else {
          const unsigned mh0 = __reduce_max_sync(0xffffffffu, h0);
          const bool p0 = h0 == mh0;
          const unsigned ml0 = __reduce_max_sync(0xffffffffu, p0 ? l0 : 0u);
          const bool q0 = p0 && l0 == ml0;
          const unsigned w00 = __ballot_sync(0xffffffffu, q0 && s0 == 0), w01 = __ballot_sync(0xffffffffu, q0);
          a0 = w00 ? __ffs((int)w00) - 1 : __ffs((int)w01) - 1 + 32;
          m0 = f64_unkey(mh0, ml0);
          const unsigned mh1 = __reduce_max_sync(0xffffffffu, h1);
          const bool p1 = h1 == mh1;
          const unsigned ml1 = __reduce_max_sync(0xffffffffu, p1 ? l1 : 0u);
          const bool q1 = p1 && l1 == ml1;
          const unsigned w10 = __ballot_sync(0xffffffffu, q1 && s1 == 0), w11 = __ballot_sync(0xffffffffu, q1);
          a1 = w10 ? __ffs((int)w10) - 1 : __ffs((int)w11) - 1 + 32;
          m1 = f64_unkey(mh1, ml1);
        }
      }
      // own-genotype substitution: v2 / v3 at index g against the maximum at index a
      VCand *xo = xmine + ((u & 1) ? 2 * Z * ROW : 0);
#pragma unroll
      for (int s = 0; s < GS; ++s) {
        const bool tx = (v2[s] > m0) || (v2[s] == m0 && g[s] < a0);
        const bool ty = (v3[s] > m1) || (v3[s] == m1 && g[s] < a1);
        VCand cx, cy;
        cx.v = tx ? v2[s] : m0; cx.i = ibase + (tx ? g[s] : a0); cx.pad = 0;
        cy.v = ty ? v3[s] : m1; cy.i = ibase + (ty ? g[s] : a1); cy.pad = 0;
        xo[32 * s] = cx;
        xo[Z * ROW + 32 * s] = cy;
      }
      __syncthreads();
      // ---- gather over the Z clusters, in source order ----
      const int slot = (tt - cs - 1) & (RV - 1);
#pragma unroll
      for (int s = 0; s < GS; ++s) {
        double gv[Z];
        int gi[Z];
#pragma unroll
        for (int w = 0; w < Z; ++w) {
          double ip;
          lds_v2f64(gadr[s][w] + ((u & 1) ? XPAR : 0u), gv[w], ip);
          gi[w] = __double2loint(ip);
        }
        const double e_u = ring.e[slot * K + j[s]];
        const double2 *rl = reinterpret_cast<const double2 *>(ring.L + ((size_t)slot * K + j[s]) * 4);
        const double2 la_u = rl[0], lb_u = rl[1];
        double best;
        int arg;
        PairTree<0, Z>::run(gv, gi, best, arg);
        const double d = on[s] ? __dadd_rn(best, e_u) : -CUDART_INF;
        v0[s] = __dadd_rn(d, la_u.x); v1[s] = __dadd_rn(d, la_u.y); v2[s] = __dadd_rn(d, lb_u.x); v3[s] = __dadd_rn(d, lb_u.y);
        if (on[s]) {
          A.psi[(size_t)tt * K + j[s]] = (unsigned char)arg;
          if (tt == last) dfin[j[s]] = d;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // first-index arg-max of the last delta (viterbiC_clonalCN.c:118)
    int arg = 0;
    for (int i = 1; i < K; ++i)
      if (dfin[arg] < dfin[i]) arg = i;
    A.path[ce - 1] = arg;            // 0-based for now; the back-trace adds 1
  }
}

// ------------------------------------------------------------------------------------------
// Back-trace (src/viterbiC_clonalCN.c:119-126: path[t-1] = psi[t][path[t]]), parallel in position.
// Row t of psi is a map state -> state, and maps compose associatively, so the pointer chase over a
// chromosome splits into tiles of R rows (tile k of a chromosome covers rows hi = ce-1-k*R down to
// lo = max(hi-R+1, cs+1)):
//   compose  one CTA per tile: thread s follows start state s through the tile -> comp[tile][s]
//   link     one thread per chromosome chains the tile maps from the arg-max end state -> entry[tile]
//   walk     one CTA per tile: re-walk from the now known entry state, coalesced store of the path
// Integer work only; identical to the sequential chase by construction.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bt_locate(const VitArgs &A, int tile, int &chr, int &lo, int &hi) {
  chr = 0;
  while (chr + 1 < A.nChr && A.tile_first[chr + 1] <= tile) ++chr;
  const int k = tile - A.tile_first[chr];
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  hi = ce - 1 - k * A.bt_rows;
  lo = max(hi - A.bt_rows + 1, cs + 1);
}

// stage psi rows lo..hi (contiguous bytes) into shared memory with 16-byte loads; returns the tile's first byte
__device__ __forceinline__ const unsigned char *bt_stage(const VitArgs &A, int lo, int hi, unsigned char *smem) {
  const size_t base = (size_t)lo * A.K, n = (size_t)(hi - lo + 1) * A.K;
  const size_t a0 = base & ~(size_t)15;
  const int shift = (int)(base - a0);
  const int chunks = (int)((shift + n + 15) >> 4);        // psi is allocated with 64 bytes of slack behind it
  const uint4 *src = reinterpret_cast<const uint4 *>(A.psi + a0);
  uint4 *dst = reinterpret_cast<uint4 *>(smem);
  for (int q = threadIdx.x; q < chunks; q += blockDim.x) dst[q] = src[q];
  __syncthreads();
  return smem + shift;
}

__global__ void __launch_bounds__(256) viterbi_bt_compose_kernel(VitArgs A) {
  extern __shared__ __align__(16) unsigned char bt_sm[];
  int chr, lo, hi;
  bt_locate(A, blockIdx.x, chr, lo, hi);
  const unsigned char *tile = bt_stage(A, lo, hi, bt_sm);
  const int K = A.K;
  for (int s = threadIdx.x; s < K; s += blockDim.x) {
    int st = s;
    for (int t = hi; t >= lo; --t) st = tile[(size_t)(t - lo) * K + st];
    A.comp[(size_t)blockIdx.x * K + s] = (unsigned char)st;
  }
}

__global__ void viterbi_bt_link_kernel(VitArgs A) {
  const int chr = blockIdx.x * blockDim.x + threadIdx.x;
  if (chr >= A.nChr) return;
  const int cs = A.chr_start[chr], ce = A.chr_start[chr + 1];
  if (cs >= ce) return;
  int st = A.path[ce - 1];             // 0-based arg-max of the last delta (forward kernel)
  A.path[ce - 1] = st + 1;
  for (int k = A.tile_first[chr]; k < A.tile_first[chr + 1]; ++k) {
    A.entry[k] = (unsigned char)st;
    st = A.comp[(size_t)k * A.K + st];
  }
}

__global__ void __launch_bounds__(256) viterbi_bt_walk_kernel(VitArgs A) {
  extern __shared__ __align__(16) unsigned char bt_sm[];
  int chr, lo, hi;
  bt_locate(A, blockIdx.x, chr, lo, hi);
  const unsigned char *tile = bt_stage(A, lo, hi, bt_sm);
  const int K = A.K, rows = hi - lo + 1;
  unsigned char *sp = bt_sm + (((size_t)A.bt_rows * K + 31) & ~(size_t)15) + 16;   // [rows] states of path[lo-1 .. hi-1]
  if (threadIdx.x == 0) {
    int st = A.entry[blockIdx.x];
    for (int t = hi; t >= lo; --t) {
      st = tile[(size_t)(t - lo) * K + st];
      sp[t - lo] = (unsigned char)st;
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < rows; q += blockDim.x) A.path[lo - 1 + q] = (int)sp[q] + 1;
}

}  // namespace titan
