// titan_fb2.cu -- forward-backward engine of the E-step on sm_100a (fp64 throughout).  See titan_fb2.h.
//
// State layout: joint state j = u + z*U (genotype fastest; R/hmmClonal.R:196, src/fwd_backC_clonalCN.c:234-235).
// One warp owns one position chunk: lane l holds the Z cluster values of genotypes l and (U > 32) l + 32, so the
// per-genotype sums are thread-local and the per-cluster sums are Z interleaved warp butterflies.
//
// Structured transition (src/fwd_backC_clonalCN.c:215-301), regrouped so that every coefficient is non-negative:
//   destination non-HET:  o[g,z] = k0*W + k1*Wz[z] + k2*Wg[g] + k3*w[g,z]
//   destination HET    :  o[g,z] = cB*W + k4*Wg[g]                          w = alpha / rowsum(source)
//
// Boundary acceptance.  A chunk that started from the vector s although its neighbour actually produced e
// (both normalised) changes the log-likelihood by sum_j gamma_j (s_j/e_j - 1) to first order and every
// downstream posterior by at most twice sum_j gamma_j |s_j/e_j - 1|, where gamma is the POSTERIOR at the
// boundary.  The forward test therefore weighs the mismatch with beta~ at the boundary (published by the
// backward chunk of the previous round) and vice versa:  sum_j w_j |s_j - e_j| <= tol * sum_j w_j e_j.
// Components the data on the other side rule out no longer force re-runs (a component-wise relative test keeps
// every such component's error alive for as long as the diagonal of the transition matrix carries it).
#include "titan_fb2.h"

#include <algorithm>

namespace titan {

// ------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kNS = 3;   // ring slots per warp
constexpr int kWPC = 4;  // warps per CTA: a warp's SM sub-partition is its index in the CTA modulo 4, so one-warp CTAs
                         // would crowd a single scheduler.  Round kernels: four independent chunks; final pass: the
                         // four quarters of one chunk.

__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_v2f64(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void bulk_s2g(void *gdst, unsigned ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// Per-warp ring of per-SNP row tiles.  A tile is TS consecutive SNP rows of up to four arrays; lane 0 issues the
// copies, every lane waits on the slot's mbarrier before reading.  `fin` adds the alpha rows and the records.
// The cursor walks the stream one position at a time (position i of the stream is SNP first + dir*i); tiles are
// fetched kNS - 1 ahead.  Everything the inner loop touches is a 32-bit shared-space address that is bumped, not
// recomputed: brow / trow (/ arow / srow) address the rows of the current position.
struct Ring {
  unsigned sbase, sbar, slot_bytes, rowB, offT, offA, offS;
  unsigned cur, brow, trow;    // current slot, its B row and its transition record
  int rstep, tstep, sstep;     // signed byte steps of the B / alpha rows, the transition records and the SNP records
  int TS, s, k, ntiles, slot_i;
  unsigned phase;
  const char *gB, *gT, *gA, *gS;             // global source of the next tile to fetch
  long long dB, dT, dS;                      // signed byte strides between consecutive tiles
  bool fin, up;
  int lane;

  __device__ __forceinline__ void issue(int sl) {      // lane 0: fetch the next tile into slot sl
    const unsigned slot = sbase + (unsigned)sl * slot_bytes, bar = sbar + 8u * (unsigned)sl;
    fence_async_smem();          // the slot was read through the generic proxy until the __syncwarp before this call
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot), "l"(gB),
                 "r"((unsigned)TS * rowB), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot + offT), "l"(gT),
                 "r"((unsigned)TS * (unsigned)sizeof(TxnRec)), "r"(bar)
                 : "memory");
    if (fin) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot + offA), "l"(gA),
                   "r"((unsigned)TS * rowB), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot + offS), "l"(gS),
                   "r"((unsigned)TS * 32u), "r"(bar)
                   : "memory");
      gA += dB;
      gS += dS;
    }
    gB += dB;
    gT += dT;
  }
  __device__ __forceinline__ void wait_slot(int sl, unsigned ph) {
    const unsigned bar = sbar + 8u * (unsigned)sl;
    unsigned ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(ph)
          : "memory");
    } while (!ok);
  }
  __device__ __forceinline__ void enter_slot() {
    cur = sbase + (unsigned)slot_i * slot_bytes;
    brow = cur + (up ? 0u : (unsigned)(TS - 1) * rowB);
    trow = cur + offT + (up ? 0u : (unsigned)(TS - 1) * (unsigned)sizeof(TxnRec));
  }
  __device__ __forceinline__ void start(const Fb2Args &A, char *smem, int TS_, bool fin_, int lane_, int dir, long long first,
                                        int npos) {
    TS = TS_; fin = fin_; lane = lane_; up = dir > 0;
    rowB = (unsigned)A.Kp * 8u;
    offT = (unsigned)TS * rowB;
    offA = offT + (unsigned)TS * (unsigned)sizeof(TxnRec);
    offS = offA + (fin ? (unsigned)TS * rowB : 0u);
    slot_bytes = offS + (fin ? (unsigned)TS * 32u : 0u);
    sbase = smem_u32(smem);
    sbar = sbase + (unsigned)kNS * slot_bytes;
    rstep = up ? (int)rowB : -(int)rowB;
    tstep = up ? (int)sizeof(TxnRec) : -(int)sizeof(TxnRec);
    sstep = up ? 32 : -32;
    ntiles = (npos + TS - 1) / TS;
    // ascending streams: tile kk holds SNP rows first + kk*TS ..; descending ones: rows first - (kk+1)*TS + 1 .. first - kk*TS;
    // posteriors are formed one SNP below the emission row of the same stream position
    const long long row0 = up ? first : first - (long long)TS + 1;
    gB = reinterpret_cast<const char *>(A.Bm + row0 * A.Kp);
    gT = reinterpret_cast<const char *>(A.txn + row0);
    gA = reinterpret_cast<const char *>(A.alpha + (row0 - 1) * A.Kp);
    gS = reinterpret_cast<const char *>(A.snp + (row0 - 1));
    const long long sgn = up ? 1 : -1;
    dB = sgn * (long long)TS * rowB;
    dT = sgn * (long long)TS * (long long)sizeof(TxnRec);
    dS = sgn * (long long)TS * 32;
    k = 0; s = 0; slot_i = 0; phase = 0u;
    if (lane == 0) {
      for (int i = 0; i < kNS; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sbar + 8u * (unsigned)i), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_async_smem();
      for (int i = 0; i < kNS && i < ntiles; ++i) issue(i);
    }
    __syncwarp();
    enter_slot();
    if (ntiles > 0) wait_slot(0, 0u);
  }
  // before the ring's shared memory is used for anything else: an mbarrier word must be invalidated first
  __device__ __forceinline__ void retire() {
    __syncwarp();
    if (lane == 0)
      for (int i = 0; i < kNS; ++i) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(sbar + 8u * (unsigned)i) : "memory");
    __syncwarp();
  }
  __device__ __forceinline__ unsigned arow() const { return brow + offA; }                       // alpha row of this position (fin)
  __device__ __forceinline__ unsigned srow() const { return cur + offS + (unsigned)(up ? s : TS - 1 - s) * 32u; }
  __device__ __forceinline__ void advance() {
    brow += rstep;
    trow += tstep;
    if (++s == TS) {
      __syncwarp();
      if (lane == 0 && k + kNS < ntiles) issue(slot_i);
      ++k;
      s = 0;
      if (++slot_i == kNS) { slot_i = 0; phase ^= 1u; }
      enter_slot();
      if (k < ntiles) wait_slot(slot_i, phase);
    }
  }
};

// Alpha rows leave through shared memory: the owned steps of a forward chunk park their rows in one of two tile
// buffers and lane 0 sends every full tile to HBM with one bulk store (cp.async.bulk.global.shared).
struct AlphaOut {
  unsigned sbuf0, sbuf1, rowB, wr;   // the two tile buffers (scalars, not an array: a dynamically indexed member would put
                                     // the whole cursor into local memory); wr: shared address of the next row to fill
  char *gdst;
  int TS, filled, which, lane;
  __device__ __forceinline__ void start(unsigned smem_addr, int TS_, unsigned rowB_, double *dst, int lane_) {
    TS = TS_; rowB = rowB_; lane = lane_;
    sbuf0 = smem_addr;
    sbuf1 = smem_addr + (unsigned)TS * rowB;
    gdst = reinterpret_cast<char *>(dst);
    filled = 0; which = 0; wr = sbuf0;
  }
  __device__ __forceinline__ void flush() {
    if (filled == 0) return;
    fence_async_smem();                         // every lane: its rows were written through the generic proxy
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(gdst, which ? sbuf1 : sbuf0, (unsigned)filled * rowB);
      bulk_commit();
      bulk_wait_read<1>();                      // the other buffer's store has finished reading it
    }
    __syncwarp();
    gdst += (size_t)filled * rowB;
    which ^= 1;
    wr = which ? sbuf1 : sbuf0;
    filled = 0;
  }
  __device__ __forceinline__ void row_done() {
    wr += rowB;
    if (++filled == TS) flush();
  }
  __device__ __forceinline__ void finish() {
    flush();
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the rows have landed, not just left shared memory
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------
// lane geometry and the structured steps
// ------------------------------------------------------------------------------------------
template <int Z, int GS>
struct LaneInfo {
  int g[GS];
  bool act[GS], het[GS];
  unsigned off[GS][Z];        // byte offset of state (g, z) inside a K-vector row (0 for inactive slots)
  unsigned hoff[GS];          // byte offset of this slot's coefficient half inside a TxnRec
  int het_lane, het_slot;     // where the diploid-HET genotype lives (-1: the table has none)
};

template <int Z, int GS>
__device__ __forceinline__ void lane_info(LaneInfo<Z, GS> &L, int U, int lane, const unsigned char *het) {
  L.het_lane = -1;
  L.het_slot = 0;
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    L.g[s] = lane + 32 * s;
    L.act[s] = L.g[s] < U;
    L.het[s] = L.act[s] && het[L.g[s]];
    L.hoff[s] = L.het[s] ? (unsigned)sizeof(TxnHalf) : 0u;
#pragma unroll
    for (int z = 0; z < Z; ++z) L.off[s][z] = L.act[s] ? 8u * (unsigned)(L.g[s] + z * U) : 0u;
  }
  for (int u = 0; u < U; ++u)
    if (het[u]) { L.het_lane = u & 31; L.het_slot = u >> 5; }
}

// forward: a <- (A_t^T applied to a) .* (B * 2^-e);  the lane's coefficient half does the HET / non-HET selection
template <int Z, int GS>
__device__ __forceinline__ void fwd_step2(double (&a)[GS][Z], const double (&B)[GS][Z], unsigned trow, const LaneInfo<Z, GS> &L,
                                          int &e) {
  double w[GS][Z], Wz[Z], Wg[GS];
  double2 c01[GS], c23[GS];
  double c4[GS];
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    const unsigned cf = trow + L.hoff[s];
    c01[s] = lds_v2f64(cf);          // irs, cW
    c23[s] = lds_v2f64(cf + 16);     // cG, cZ
    c4[s] = lds_f64(cf + 32);        // cw
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      w[s][z] = a[s][z] * c01[s].x;
      Wz[z] = (s == 0) ? w[s][z] : Wz[z] + w[s][z];
      Wg[s] = (z == 0) ? w[s][z] : Wg[s] + w[s][z];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)      // Z interleaved butterflies: Z independent shuffles per stage
#pragma unroll
    for (int z = 0; z < Z; ++z) Wz[z] += __shfl_xor_sync(0xffffffffu, Wz[z], o);
  double W = Wz[0];
#pragma unroll
  for (int z = 1; z < Z; ++z) W += Wz[z];
  const double sc = pow2_rescale(W, e);
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    const double base = c01[s].y * W + c23[s].x * Wg[s];
#pragma unroll
    for (int z = 0; z < Z; ++z) a[s][z] = (base + c23[s].y * Wz[z] + c4[s] * w[s][z]) * (B[s][z] * sc);
  }
}

// backward: b <- A_{t+1} applied to (b .* B_{t+1}), rescaled by a power of two.  The table holds at most one
// diploid-HET genotype, so the HET mass is a broadcast from its lane instead of a butterfly.
template <int Z, int GS>
__device__ __forceinline__ void bwd_step2(double (&b)[GS][Z], const double (&B)[GS][Z], unsigned trow, const LaneInfo<Z, GS> &L,
                                          int &e) {
  double bb[GS][Z], loc[GS], red[Z];
#pragma unroll
  for (int s = 0; s < GS; ++s) {
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      bb[s][z] = b[s][z] * B[s][z];
      loc[s] = (z == 0) ? bb[s][z] : loc[s] + bb[s][z];
      const double nh = L.het[s] ? 0.0 : bb[s][z];
      red[z] = (s == 0) ? nh : red[z] + nh;
    }
  }
  double H = 0.0;
  if (L.het_lane >= 0) {
    double hl = loc[0];
#pragma unroll
    for (int s = 1; s < GS; ++s) hl = (L.het_slot == s) ? loc[s] : hl;
    H = __shfl_sync(0xffffffffu, hl, L.het_lane);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int z = 0; z < Z; ++z) red[z] += __shfl_xor_sync(0xffffffffu, red[z], o);
  double Wn = red[0];
#pragma unroll
  for (int z = 1; z < Z; ++z) Wn += red[z];
  const double sc = pow2_rescale(Wn + H, e);
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    const unsigned cf = trow + L.hoff[s];
    const double2 c01 = lds_v2f64(cf), c23 = lds_v2f64(cf + 16), c45 = lds_v2f64(cf + 32), c67 = lds_v2f64(cf + 48);
    // TxnHalf: irs, cW | cG, cZ | cw, k0 | cB, k1
    const double irs = L.act[s] ? c01.x * sc : 0.0;
    const double base = c45.y * Wn + c67.x * H + c23.x * loc[s];
#pragma unroll
    for (int z = 0; z < Z; ++z) b[s][z] = (base + c67.y * red[z] + c45.x * bb[s][z]) * irs;
  }
}

template <int Z, int GS>
__device__ __forceinline__ void load_row(double (&v)[GS][Z], const void *row, const LaneInfo<Z, GS> &L) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) v[s][z] = L.act[s] ? *reinterpret_cast<const double *>(static_cast<const char *>(row) + L.off[s][z]) : 0.0;
}

// the same through the read-only path (ld.global.nc): only for vectors no thread of the running kernel writes; lets
// the compiler hoist the loads above unrelated global stores
template <int Z, int GS>
__device__ __forceinline__ void load_row_nc(double (&v)[GS][Z], const void *row, const LaneInfo<Z, GS> &L) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) v[s][z] = L.act[s] ? __ldg(reinterpret_cast<const double *>(static_cast<const char *>(row) + L.off[s][z])) : 0.0;
}

template <int Z, int GS>
__device__ __forceinline__ void load_row_s(double (&v)[GS][Z], unsigned srow, const LaneInfo<Z, GS> &L) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) v[s][z] = L.act[s] ? lds_f64(srow + L.off[s][z]) : 0.0;
}

template <int Z, int GS>
__device__ __forceinline__ void store_row_s(unsigned srow, const double (&v)[GS][Z], const LaneInfo<Z, GS> &L) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
    if (L.act[s])
#pragma unroll
      for (int z = 0; z < Z; ++z) sts_f64(srow + L.off[s][z], v[s][z]);
}

template <int Z, int GS>
__device__ __forceinline__ void store_row(void *row, const double (&v)[GS][Z], const LaneInfo<Z, GS> &L, double scale) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
    if (L.act[s])
#pragma unroll
      for (int z = 0; z < Z; ++z) *reinterpret_cast<double *>(static_cast<char *>(row) + L.off[s][z]) = v[s][z] * scale;
}

template <int Z, int GS>
__device__ __forceinline__ void store_row(void *row, const double (&v)[GS][Z], const LaneInfo<Z, GS> &L) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
    if (L.act[s])
#pragma unroll
      for (int z = 0; z < Z; ++z) *reinterpret_cast<double *>(static_cast<char *>(row) + L.off[s][z]) = v[s][z];
}

template <int Z, int GS>
__device__ __forceinline__ double vec_sum(const double (&v)[GS][Z]) {
  double loc = 0.0;
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) loc += v[s][z];
  return warp_sum(loc);
}

template <int Z, int GS>
__device__ __forceinline__ void scale_vec(double (&v)[GS][Z], double f) {
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) v[s][z] *= f;
}

// posterior-weighted comparison of the boundary vector a chunk used with the one its neighbour produced;
// `want` is left in v[][] (the vector a re-run starts from)
template <int Z, int GS>
__device__ __forceinline__ bool boundary_ok(double (&v)[GS][Z], const double *used, const double *want, const double *wgt,
                                            const LaneInfo<Z, GS> &L, double tol) {
  double u[GS][Z], p[GS][Z];
  load_row<Z, GS>(u, used, L);
  load_row<Z, GS>(v, want, L);
  load_row<Z, GS>(p, wgt, L);
  double num = 0.0, den = 0.0;
  bool bad = false;
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      num += p[s][z] * fabs(u[s][z] - v[s][z]);
      den += p[s][z] * v[s][z];
      bad |= (u[s][z] != u[s][z]) || (v[s][z] != v[s][z]);
    }
  num = warp_sum(num);
  den = warp_sum(den);
  const bool ok = (num <= tol * den) && (den > 0.0);      // false for NaN as well
  return ok && !__any_sync(0xffffffffu, bad);
}

template <int Z, int GS>
__device__ __forceinline__ void copy_vec(double *dst, const double *src, const LaneInfo<Z, GS> &L) {
  double v[GS][Z];
  load_row<Z, GS>(v, src, L);
  store_row<Z, GS>(dst, v, L);
}

// one forward step at the cursor's position; OWNED steps park alpha~ for the bulk store and account for the log-likelihood
template <int Z, int GS, bool OWNED>
__device__ __forceinline__ void fwd_one(double (&a)[GS][Z], Ring &ring, const LaneInfo<Z, GS> &L, unsigned mOff, double &acc_m,
                                        int &acc_e, AlphaOut &out) {
  double B[GS][Z];
  load_row_s<Z, GS>(B, ring.brow, L);
  int e;
  fwd_step2<Z, GS>(a, B, ring.trow, L, e);
  if (OWNED) {
    acc_m += lds_f64(ring.brow + mOff);
    acc_e += e;
    store_row_s<Z, GS>(out.wr, a, L);
    out.row_done();
  }
  ring.advance();
}

// ------------------------------------------------------------------------------------------
// forward chunk (reference recursion: src/fwd_backC_clonalCN.c:111-142, here scaled linear space)
// ------------------------------------------------------------------------------------------
template <int Z, int GS>
__device__ __forceinline__ void fb2_forward(const Fb2Args &A, int c, int lane, char *smem, int TS) {
  const Chunk ch = A.chunks[c];
  const int U = A.U, K = A.K;
  LaneInfo<Z, GS> L;
  lane_info<Z, GS>(L, U, lane, A.het);
  const size_t cur = (size_t)(A.round & 1) * A.nChunks * K, prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  double *startv = A.startF + (size_t)c * K;
  double a[GS][Z];
  int tb;
  bool have_start;
  if (A.round > 0) {
    bool skip = (ch.t0 == ch.cs);       // chromosome-leading chunks start from pi: exact in round 0
    if (A.use_solved) {                 // probe mode: the probe kernel tested the boundary, the chain solve predicted the start
      skip = skip || !A.solvedF[c];
      if (!skip) load_row<Z, GS>(a, A.solF + (size_t)c * K, L);
    } else if (!skip) {
      skip = boundary_ok<Z, GS>(a, startv, A.endF + prv + (size_t)(c - 1) * K, A.extB + prv + (size_t)c * K, L, A.tol);
    }
    if (skip) {
      copy_vec<Z, GS>(A.endF + cur + (size_t)c * K, A.endF + prv + (size_t)c * K, L);
      copy_vec<Z, GS>(A.extF + cur + (size_t)c * K, A.extF + prv + (size_t)c * K, L);
      return;
    }
    tb = ch.t0;
    have_start = true;
  } else {
    tb = max(ch.t0 - A.warm, ch.cs);
    have_start = false;
  }
  if (lane == 0 && !(A.use_solved && A.round > 0)) atomicAdd(A.rerun + A.round, 1u);
  if (lane == 0 && A.round > 0) atomicAdd(A.ver + c, 1u);      // a final pass of this chunk started earlier is stale now

  const int n_own = ch.t1 - tb;                 // stream positions q < n_own are SNPs tb + q of this chunk
  const bool has_ext = ch.t1 < ch.ce;           // one more step: alpha^ at the first SNP of the next chunk
  const int q0 = ch.t0 - tb;
  Ring ring;
  ring.start(A, smem, TS, false, lane, +1, tb, n_own + (has_ext ? 1 : 0));
  AlphaOut out;
  out.start(ring.sbar + 8u * kNS + 8u, TS, ring.rowB, A.alpha + (size_t)ch.t0 * A.Kp, lane);
  const unsigned mOff = 8u * (unsigned)K;

  double acc_m = 0.0;        // sum of emission shifts over the owned range
  int acc_e = 0;             // sum of power-of-two rescale exponents over the owned range
  double s_start = 1.0;      // mass of the vector entering the owned range
  int q = 0;
  if (!have_start) {
    // stream position 0 has no predecessor vector: chromosome start (pi) or a flat speculative start
    double B[GS][Z];
    load_row_s<Z, GS>(B, ring.brow, L);
    if (tb == ch.cs) {
      load_row<Z, GS>(a, A.pi, L);
#pragma unroll
      for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int z = 0; z < Z; ++z) a[g][z] *= B[g][z];
    } else {
#pragma unroll
      for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int z = 0; z < Z; ++z) a[g][z] = B[g][z];
    }
    if (q0 == 0) {           // only a chromosome-leading chunk owns its very first position
      acc_m += lds_f64(ring.brow + mOff);
      store_row_s<Z, GS>(out.wr, a, L);
      out.row_done();
    }
    ring.advance();
    q = 1;
    for (; q < q0; ++q) fwd_one<Z, GS, false>(a, ring, L, mOff, acc_m, acc_e, out);      // warm-up: nothing stored
  }
  if (ch.t0 != ch.cs) {      // entering the owned range: record / publish alpha~_{t0-1}
    s_start = vec_sum<Z, GS>(a);
    if (!have_start) { scale_vec<Z, GS>(a, 1.0 / s_start); s_start = 1.0; }
    store_row<Z, GS>(startv, a, L);
  }
  for (; q < n_own; ++q) fwd_one<Z, GS, true>(a, ring, L, mOff, acc_m, acc_e, out);
  out.finish();
  const double s_end = vec_sum<Z, GS>(a);
  store_row<Z, GS>(A.endF + cur + (size_t)c * K, a, L, 1.0 / s_end);
  if (lane == 0) A.chunk_ll[c] = (log(s_end) - log(s_start)) + acc_m + (double)acc_e * 0.6931471805599453094;
  if (has_ext) {
    fwd_one<Z, GS, false>(a, ring, L, mOff, acc_m, acc_e, out);
    store_row<Z, GS>(A.extF + cur + (size_t)c * K, a, L, 1.0 / vec_sum<Z, GS>(a));
  }
}

template <int Z, int GS>
__device__ __forceinline__ void bwd_one(double (&b)[GS][Z], Ring &ring, const LaneInfo<Z, GS> &L) {
  double B[GS][Z];
  load_row_s<Z, GS>(B, ring.brow, L);
  int e;
  bwd_step2<Z, GS>(b, B, ring.trow, L, e);
  ring.advance();
}

// ------------------------------------------------------------------------------------------
// backward chunk, boundary rounds: the beta recursion alone (src/fwd_backC_clonalCN.c:148-175)
// b[] holds beta~_{tn}; tn == ce means "past the end" (beta_{ce-1} = 1).  Stream position j is SNP tn - j:
// step j turns beta~_{tn-j} into beta~_{tn-j-1} with B and the transition record of SNP tn - j.
// The vectors entering the lower three quarters of the chunk are parked for the final pass (subB).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int quarter_start(const Chunk &ch, int k) { return ch.t0 + (int)(((long long)(ch.t1 - ch.t0) * k) / kWPC); }

template <int Z, int GS>
__device__ __forceinline__ void fb2_backward(const Fb2Args &A, int c, int lane, char *smem, int TS) {
  const Chunk ch = A.chunks[c];
  const int U = A.U, K = A.K;
  LaneInfo<Z, GS> L;
  lane_info<Z, GS>(L, U, lane, A.het);
  const size_t cur = (size_t)(A.round & 1) * A.nChunks * K, prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  double *startv = A.startB + (size_t)c * K;
  double b[GS][Z];
  int tn;
  bool from_boundary;
  if (A.round > 0) {
    bool skip = (ch.t1 == ch.ce);
    if (A.use_solved) {
      skip = skip || !A.solvedB[c];
      if (!skip) load_row<Z, GS>(b, A.solB + (size_t)c * K, L);
    } else if (!skip) {
      skip = boundary_ok<Z, GS>(b, startv, A.endB + prv + (size_t)(c + 1) * K, A.extF + prv + (size_t)c * K, L, A.tol);
    }
    if (skip) {
      copy_vec<Z, GS>(A.endB + cur + (size_t)c * K, A.endB + prv + (size_t)c * K, L);
      copy_vec<Z, GS>(A.extB + cur + (size_t)c * K, A.extB + prv + (size_t)c * K, L);
      return;
    }
    tn = ch.t1;
    from_boundary = true;
  } else {
    tn = min(ch.t1 + A.warm, ch.ce);
    from_boundary = false;
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int z = 0; z < Z; ++z) b[s][z] = L.act[s] ? 1.0 : 0.0;
  }
  if (lane == 0 && !(A.use_solved && A.round > 0)) atomicAdd(A.rerun + A.round, 1u);
  if (lane == 0 && A.round > 0) atomicAdd(A.ver + c, 1u);

  const int n_main = tn - ch.t0;                // steps j < n_main produce beta~ at SNPs tn-1 .. t0
  const bool has_ext = ch.t0 > ch.cs;           // one more step: beta~ at the last SNP of the previous chunk
  Ring ring;
  ring.start(A, smem, TS, false, lane, -1, tn, n_main + (has_ext ? 1 : 0));
  int j = 0;
  if (tn == ch.ce) {                            // position 0 is past the end: beta_{ce-1} = 1, no step
    ring.advance();
    j = 1;
  }
  for (; j < tn - ch.t1; ++j) bwd_one<Z, GS>(b, ring, L);       // warm-up down to beta~_{t1}
  if (ch.t1 < ch.ce) {                          // b[] is beta~_{t1}: normalise what the warm-up produced, publish it
    if (!from_boundary) scale_vec<Z, GS>(b, 1.0 / vec_sum<Z, GS>(b));
    store_row<Z, GS>(startv, b, L);
  }
  // quarters kWPC-2 .. 0 of the chunk start from beta~ at the first SNP of the quarter above them
  for (int k = kWPC - 1; k >= 1; --k) {
    const int jq = tn - quarter_start(ch, k);   // after jq steps b[] is beta~ at SNP quarter_start(k)
    for (; j < jq; ++j) bwd_one<Z, GS>(b, ring, L);
    store_row<Z, GS>(A.subB + ((size_t)c * (kWPC - 1) + (k - 1)) * K, b, L);
  }
  for (; j < n_main; ++j) bwd_one<Z, GS>(b, ring, L);
  store_row<Z, GS>(A.endB + cur + (size_t)c * K, b, L, 1.0 / vec_sum<Z, GS>(b));
  if (has_ext) {
    bwd_one<Z, GS>(b, ring, L);
    store_row<Z, GS>(A.extB + cur + (size_t)c * K, b, L, 1.0 / vec_sum<Z, GS>(b));
  }
}

// one launch per round: warps [0, nChunks) run forward chunks, [nChunks, 2 nChunks) backward chunks
template <int Z, int GS>
__global__ void __launch_bounds__(32 * kWPC, 3) fb2_round_kernel(Fb2Args A, int TS, int ring_stride) {
  extern __shared__ __align__(128) char fb2_smem[];
  if (A.use_solved) {
    if (A.round > 0 && A.rerun[A.round] == 0) return;        // probe mode: this cycle's boundary test found nothing to re-run
  } else if (A.round > 1 && A.rerun[A.round - 1] == 0) return;   // fixed point already reached (queued round)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int id = blockIdx.x * kWPC + warp;                   // the warps of a CTA are independent of one another
  char *smem = fb2_smem + (size_t)warp * ring_stride;
  if (id < A.nChunks) fb2_forward<Z, GS>(A, id, lane, smem, TS);
  else if (id < 2 * A.nChunks) fb2_backward<Z, GS>(A, id - A.nChunks, lane, smem, TS);
}

// ------------------------------------------------------------------------------------------
// Probe-accelerated rounds.  Re-running a dirty chunk from its neighbour's latest vector moves information by one
// chunk per round, so a run of m dirty chunks costs m rounds of one chunk latency each.  The chunk map is linear,
// and what a boundary vector has to get right is small: the posterior at a boundary lives on a handful of states
// (competing genotypes, the same genotype in two clusters).  So each dirty chunk propagates three pieces of the
// vector it started from -- the two states of largest posterior weight as unit vectors, and the remainder --
// on three otherwise idle warps (fb2_probe_kernel), a per-chromosome pass chains the resulting 3-dimensional maps
// through every run of dirty chunks (fb2_solve_kernel), and the round kernel re-runs the dirty chunks once from
// the predicted vectors.  The next boundary test decides; nothing is accepted on the strength of a prediction.
// ------------------------------------------------------------------------------------------


// the kProbes - 1 states of largest posterior weight want_j * wgt_j (ties: lowest index), identical in every warp that asks
template <int Z, int GS>
__device__ __forceinline__ void top_states(const double (&want)[GS][Z], const double *wgt, const LaneInfo<Z, GS> &L, int U,
                                           int (&found)[kProbes - 1]) {
  double g[GS][Z];
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) g[s][z] = L.act[s] ? want[s][z] * wgt[L.g[s] + z * U] : -1.0;
#pragma unroll
  for (int pass = 0; pass < kProbes - 1; ++pass) found[pass] = -1;
#pragma unroll
  for (int pass = 0; pass < kProbes - 1; ++pass) {
    double bv = -2.0;
    int bj = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        const int j = L.g[s] + z * U;
        bool taken = false;
#pragma unroll
        for (int q = 0; q < kProbes - 1; ++q) taken |= (q < pass) && (j == found[q]);
        const double v = (L.act[s] && !taken) ? g[s][z] : -2.0;
        if (v > bv || (v == bv && j < bj)) { bv = v; bj = j; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
    }
    found[pass] = bj;
  }
}


template <int Z, int GS>
__device__ __forceinline__ void fb2_probe(const Fb2Args &A, int c, int p, bool backward, int lane, char *smem, int TS) {
  const Chunk ch = A.chunks[c];
  const int U = A.U, K = A.K;
  LaneInfo<Z, GS> L;
  lane_info<Z, GS>(L, U, lane, A.het);
  const size_t prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  unsigned char *dirty = backward ? A.dirtyB : A.dirtyF;
  if (backward ? (ch.t1 == ch.ce) : (ch.t0 == ch.cs)) {      // no neighbour on that side: exact since round 0
    if (p == 0 && lane == 0) { dirty[c] = 0; (backward ? A.probedB : A.probedF)[c] = 0; }
    return;
  }
  const double *used = (backward ? A.startB : A.startF) + (size_t)c * K;
  const double *wantp = backward ? A.endB + prv + (size_t)(c + 1) * K : A.endF + prv + (size_t)(c - 1) * K;
  const double *wgt = (backward ? A.extF : A.extB) + prv + (size_t)c * K;
  double v[GS][Z];
  const bool ok = boundary_ok<Z, GS>(v, used, wantp, wgt, L, A.tol);
  // a clean chunk is probed as well when one of the probe_window chunks before it (in recursion order) is dirty:
  // that chunk's re-run will shift this chunk's input, and the chain solve can then carry the prediction on
  bool active = !ok;
  if (!active && A.probe_window > 0) {
    double tmp[GS][Z];
    for (int u = 1; u <= A.probe_window && !active; ++u) {
      const int cu = backward ? c + u : c - u;
      if (cu < 0 || cu >= A.nChunks) break;
      const Chunk cc = A.chunks[cu];
      if (cc.chr != ch.chr || (backward ? (cc.t1 == cc.ce) : (cc.t0 == cc.cs))) break;
      const double *u_used = (backward ? A.startB : A.startF) + (size_t)cu * K;
      const double *u_want = backward ? A.endB + prv + (size_t)(cu + 1) * K : A.endF + prv + (size_t)(cu - 1) * K;
      const double *u_wgt = (backward ? A.extF : A.extB) + prv + (size_t)cu * K;
      active = !boundary_ok<Z, GS>(tmp, u_used, u_want, u_wgt, L, A.tol);
    }
  }
  if (p == 0 && lane == 0) {
    dirty[c] = ok ? 0 : 1;
    (backward ? A.probedB : A.probedF)[c] = active ? 1 : 0;
    if (!ok) atomicAdd(A.rerun + A.round, 1u);
  }
  if (!active) return;
  int jt[kProbes - 1];
  top_states<Z, GS>(v, wgt, L, U, jt);
  if (p == 0 && lane == 0) {
    int *J = (backward ? A.probeJB : A.probeJF) + (kProbes - 1) * c;
#pragma unroll
    for (int q = 0; q < kProbes - 1; ++q) J[q] = jt[q];
  }
  // piece p of the vector this chunk last started from
  load_row<Z, GS>(v, used, L);
#pragma unroll
  for (int s = 0; s < GS; ++s)
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      const int j = L.g[s] + z * U;
      bool top = false;
      int mine = -1;
#pragma unroll
      for (int q = 0; q < kProbes - 1; ++q) {
        top |= (j == jt[q]);
        mine = (q == p) ? jt[q] : mine;
      }
      double x;
      if (p < kProbes - 1) x = (j == mine) ? 1.0 : 0.0;
      else x = top ? 0.0 : v[s][z];
      v[s][z] = L.act[s] ? x : 0.0;
    }
  long long E = 0;
  const int n = ch.t1 - ch.t0;
  Ring ring;
  ring.start(A, smem, TS, false, lane, backward ? -1 : +1, backward ? ch.t1 : ch.t0, n);
  for (int q = 0; q < n; ++q) {
    double B[GS][Z];
    load_row_s<Z, GS>(B, ring.brow, L);
    int e;
    if (backward) bwd_step2<Z, GS>(v, B, ring.trow, L, e);
    else fwd_step2<Z, GS>(v, B, ring.trow, L, e);
    E += e;
    ring.advance();
  }
  store_row<Z, GS>((backward ? A.probeB : A.probeF) + ((size_t)c * kProbes + p) * K, v, L);
  if (lane == 0) ((backward ? A.probeEB : A.probeEF) + (size_t)c * kProbes)[p] = (int)max(-2000000000LL, min(2000000000LL, E));
}

// warp id -> (direction, chunk, probe); clean chunks cost one boundary test
template <int Z, int GS>
__global__ void __launch_bounds__(32 * kWPC, 3) fb2_probe_kernel(Fb2Args A, int TS, int ring_stride) {
  extern __shared__ __align__(128) char fb2_smem[];
  if (A.round > 1 && A.rerun[A.round - 1] == 0) return;      // the previous test already found every boundary clean
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int id = blockIdx.x * kWPC + warp;
  if (id >= 2 * A.nChunks * kProbes) return;
  const int p = id % kProbes, cd = id / kProbes;
  const bool backward = cd >= A.nChunks;
  fb2_probe<Z, GS>(A, backward ? cd - A.nChunks : cd, p, backward, lane, fb2_smem + (size_t)warp * ring_stride, TS);
}

// One warp per (chromosome, direction): walk the chunks in recursion order carrying the best available boundary
// vector; for a dirty chunk express that vector in the chunk's three probe pieces and predict the chunk's end.
template <int Z, int GS>
__global__ void __launch_bounds__(32 * kWPC) fb2_solve_kernel(Fb2Args A) {
  if (A.rerun[A.round] == 0) return;
  const int lane = threadIdx.x & 31;
  const int id = blockIdx.x * kWPC + (threadIdx.x >> 5);
  if (id >= 2 * A.nChr) return;
  const bool backward = id >= A.nChr;
  const int chr = backward ? id - A.nChr : id;
  const int U = A.U, K = A.K;
  LaneInfo<Z, GS> L;
  lane_info<Z, GS>(L, U, lane, A.het);
  const size_t prv = (size_t)((A.round & 1) ^ 1) * A.nChunks * K;
  const int c_lo = A.chr_chunk0[chr], c_hi = A.chr_chunk0[chr + 1];     // chunks [c_lo, c_hi)
  const unsigned char *dirty = backward ? A.dirtyB : A.dirtyF;
  unsigned char *solved = backward ? A.solvedB : A.solvedF;
  const unsigned char *probed = backward ? A.probedB : A.probedF;
  const double *endv = (backward ? A.endB : A.endF) + prv;
  const double *startv = backward ? A.startB : A.startF;
  const double *probe = backward ? A.probeB : A.probeF;
  const int *probeE = backward ? A.probeEB : A.probeEF;
  const int *probeJ = backward ? A.probeJB : A.probeJF;
  double *sol = backward ? A.solB : A.solF;
  double cur[GS][Z];
  bool have = false;              // cur[] holds a PREDICTED end vector of the previous chunk in recursion order
  const int nC = c_hi - c_lo;
  for (int i = 0; i < nC; ++i) {
    const int c = backward ? c_hi - 1 - i : c_lo + i;
    bool run = (i > 0) && dirty[c];
    if (i > 0 && !run && have && probed[c]) {
      // clean against its neighbour's current end, but that neighbour is about to be re-run: test the prediction
      double tmp[GS][Z];
      const double *wgt = (backward ? A.extF : A.extB) + prv + (size_t)c * K;
      store_row<Z, GS>(sol + (size_t)c * K, cur, L);
      __syncwarp();
      run = !boundary_ok<Z, GS>(tmp, startv + (size_t)c * K, sol + (size_t)c * K, wgt, L, A.tol);
    }
    if (!run) {                   // first chunk in recursion order, or clean: keeps its own end
      if (lane == 0) solved[c] = 0;
      have = false;
      continue;
    }
    if (!have) load_row_nc<Z, GS>(cur, endv + (size_t)(backward ? c + 1 : c - 1) * K, L);     // the neighbour's actual end
    store_row<Z, GS>(sol + (size_t)c * K, cur, L);
    if (lane == 0) solved[c] = 1;
    // coefficients of cur[] in the three pieces of the vector the chunk started from
    int jt[kProbes - 1];
#pragma unroll
    for (int q = 0; q < kProbes - 1; ++q) jt[q] = __ldg(probeJ + (kProbes - 1) * c + q);
    // everything this chunk's prediction reads besides cur[]: issued together, ahead of the reductions
    int pe[kProbes];
    double pq[kProbes][GS][Z], old[GS][Z];
#pragma unroll
    for (int q = 0; q < kProbes; ++q) {
      pe[q] = __ldg(probeE + c * kProbes + q);
      load_row_nc<Z, GS>(pq[q], probe + ((size_t)c * kProbes + q) * K, L);
    }
    load_row_nc<Z, GS>(old, startv + (size_t)c * K, L);
    double x[kProbes];            // coefficients: the probed states' components, then the remainder's mass ratio
    double rn = 0.0, ro = 0.0;
#pragma unroll
    for (int q = 0; q < kProbes; ++q) x[q] = 0.0;
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        const int j = L.g[s] + z * U;
        const double v = cur[s][z], o = old[s][z];
        bool top = false;
#pragma unroll
        for (int q = 0; q < kProbes - 1; ++q) {
          x[q] += (j == jt[q]) ? v : 0.0;
          top |= (j == jt[q]);
        }
        const bool rest = L.act[s] && !top;
        rn += rest ? v : 0.0;
        ro += rest ? o : 0.0;
      }
#pragma unroll
    for (int q = 0; q < kProbes - 1; ++q) x[q] = warp_sum(x[q]);
    rn = warp_sum(rn); ro = warp_sum(ro);
    x[kProbes - 1] = ro > 0.0 ? rn / ro : 0.0;
    int Em = -2147483647;
#pragma unroll
    for (int q = 0; q < kProbes; ++q)
      if (x[q] > 0.0) Em = max(Em, pe[q]);
    double e[GS][Z];
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int z = 0; z < Z; ++z) e[s][z] = 0.0;
#pragma unroll
    for (int q = 0; q < kProbes; ++q) {
      const double f = x[q] > 0.0 ? scalbn(x[q], max(pe[q] - Em, -2000)) : 0.0;
#pragma unroll
      for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int z = 0; z < Z; ++z) e[s][z] += f * pq[q][s][z];
    }
    const double tot = vec_sum<Z, GS>(e);
    have = (tot > 0.0) && (tot < CUDART_INF);
    if (have) {
      const double inv = 1.0 / tot;
#pragma unroll
      for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int z = 0; z < Z; ++z) cur[s][z] = e[s][z] * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// final pass: beta from the verified boundaries, posteriors, marginals, the six weighted sums.
// One CTA per chunk, one warp per quarter of the chunk (the backward chunk parked beta~ at the quarter
// boundaries), partial sums combined in shared memory in a fixed order.
// reference: src/fwd_backC_clonalCN.c:148-192; R/hmmClonal.R:230-234; R/paramEstimation.R:34-40
// ------------------------------------------------------------------------------------------
template <int Z, int GS, bool WRITE_POST, bool FROM_PY>
__global__ void __launch_bounds__(32 * kWPC, 3) fb2_final_kernel(Fb2Args A, int TS, int ring_stride) {
  extern __shared__ __align__(128) char fb2_smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.x;
  unsigned int ver0 = 0;
  if (A.final_mode != 0) {                    // uniform per CTA
    ver0 = A.ver[c];
    if (A.fin_ver[c] == ver0) return;                                                              // an earlier pass is still valid
    if (A.final_mode == 1 && A.rerun[A.round] != 0 && (A.solvedF[c] | A.solvedB[c])) return;     // re-run in flight right now
  }
  const Chunk ch = A.chunks[c];
  const int U = A.U, K = A.K;
  LaneInfo<Z, GS> L;
  lane_info<Z, GS>(L, U, lane, A.het);
  const int ta = quarter_start(ch, warp), tn = quarter_start(ch, warp + 1);      // this warp owns SNPs [ta, tn)
  const int n = tn - ta;
  double st[6][GS][Z];
#pragma unroll
  for (int qq = 0; qq < 6; ++qq)
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int z = 0; z < Z; ++z) st[qq][s][z] = 0.0;
  if (n > 0) {
    double b[GS][Z];
    if (warp < kWPC - 1) load_row<Z, GS>(b, A.subB + ((size_t)c * (kWPC - 1) + warp) * K, L);
    else if (ch.t1 < ch.ce) load_row<Z, GS>(b, A.startB + (size_t)c * K, L);
    else {
#pragma unroll
      for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int z = 0; z < Z; ++z) b[s][z] = L.act[s] ? 1.0 : 0.0;
    }
    Ring ring;
    ring.start(A, fb2_smem_all + (size_t)warp * ring_stride, TS, true, lane, -1, tn, n);
    for (int j = 0; j < n; ++j) {
      const int t = tn - j - 1;                 // the SNP whose posterior this step forms
      if (tn - j < ch.ce) {
        double B[GS][Z];
        load_row_s<Z, GS>(B, ring.brow, L);
        int e;
        bwd_step2<Z, GS>(b, B, ring.trow, L, e);
      }
      double p[GS][Z];
      load_row_s<Z, GS>(p, ring.arow(), L);
      double loc = 0.0;
#pragma unroll
      for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int z = 0; z < Z; ++z) { p[g][z] *= b[g][z]; loc += p[g][z]; }
      const double inv = 1.0 / warp_sum(loc);
      scale_vec<Z, GS>(p, inv);
      if (!FROM_PY) {
        const unsigned sr = ring.srow();
        const double2 r01 = lds_v2f64(sr), r23 = lds_v2f64(sr + 16);
        SnpRec rec;
        rec.x = r01.x; rec.nx = r01.y; rec.lr = r23.x; rec.lb = r23.y;
        const double l2 = rec.lr * rec.lr;
#pragma unroll
        for (int g = 0; g < GS; ++g)
#pragma unroll
          for (int z = 0; z < Z; ++z) {
            st[0][g][z] += rec.x * p[g][z];
            st[1][g][z] += rec.nx * p[g][z];
            st[2][g][z] += rec.lr * p[g][z];
            st[3][g][z] += rec.lb * p[g][z];
            st[4][g][z] += p[g][z];
            st[5][g][z] += l2 * p[g][z];
          }
      } else if (A.loggamma != nullptr) {
        double lg[GS][Z];
#pragma unroll
        for (int g = 0; g < GS; ++g)
#pragma unroll
          for (int z = 0; z < Z; ++z) lg[g][z] = log(p[g][z]);
        store_row<Z, GS>(A.loggamma + (size_t)t * K, lg, L);
      }
      const bool first = (t == ch.cs);
      if (WRITE_POST || first) {
        double rz[Z];
#pragma unroll
        for (int z = 0; z < Z; ++z) {
          double v = 0.0;
#pragma unroll
          for (int g = 0; g < GS; ++g) v += p[g][z];
          rz[z] = warp_sum(v);
        }
        double *dG = nullptr, *dZ = nullptr;
        if (WRITE_POST && A.rhoG != nullptr) { dG = A.rhoG + (size_t)t * U; dZ = A.rhoZ + (size_t)t * Z; }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 1) {
            if (!(first && A.rhoG0 != nullptr)) break;
            dG = A.rhoG0 + (size_t)ch.chr * U;
            dZ = A.rhoZ0 + (size_t)ch.chr * Z;
          }
          if (dG == nullptr) continue;
#pragma unroll
          for (int g = 0; g < GS; ++g) {
            double gs = 0.0;
#pragma unroll
            for (int z = 0; z < Z; ++z) gs += p[g][z];
            if (L.act[g]) dG[L.g[g]] = gs;
          }
#pragma unroll
          for (int z = 0; z < Z; ++z)
            if (lane == z) dZ[z] = rz[z];
        }
      }
      ring.advance();
    }
    ring.retire();
  }
  if (!FROM_PY) {
    // combine the four quarters in a fixed order: quarter sums to shared memory, warp 0 adds them up
    __syncthreads();                              // every ring of the CTA has been drained and its mbarriers invalidated
    double *sm = reinterpret_cast<double *>(fb2_smem_all + (size_t)kWPC * ring_stride);
    double *mine = sm + (size_t)warp * 6 * K;
#pragma unroll
    for (int qq = 0; qq < 6; ++qq) store_row<Z, GS>(mine + qq * K, st[qq], L);
    __syncthreads();
    if (warp == 0) {
      double *dst = A.part + (size_t)c * 6 * K;
#pragma unroll
      for (int qq = 0; qq < 6; ++qq) {
        double acc[GS][Z], v[GS][Z];
        load_row<Z, GS>(acc, sm + qq * K, L);
#pragma unroll
        for (int w = 1; w < kWPC; ++w) {
          load_row<Z, GS>(v, sm + ((size_t)w * 6 + qq) * K, L);
#pragma unroll
          for (int g = 0; g < GS; ++g)
#pragma unroll
            for (int z = 0; z < Z; ++z) acc[g][z] += v[g][z];
        }
        store_row<Z, GS>(dst + qq * K, acc, L);
      }
      if (A.final_mode != 0 && lane == 0) A.fin_ver[c] = ver0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// emission: B[j,t] = exp(py[j,t] - m_t), m_t = max_j py[j,t] rounded to float (any common shift is exact once
// it is added back to the log-likelihood).  One pass, one warp per run of SNPs, no dependence between SNPs.
// reference: R/hmmClonal.R:491-523,617-634 (densities and state loops), :160-171 (combine)
// ------------------------------------------------------------------------------------------
constexpr int kEmRun = 32;

// SRC: 0 binomial allele model, 1 materialised log emissions (legacy .Call path), 2 Gaussian allele model
// (bivariate normal on (ref / tumDepth, logR), R/hmmClonal.R:525-559)
template <int Z, int GS, int SRC>
__global__ void __launch_bounds__(128) fb2_emission_kernel(EmArgs E) {
  constexpr bool FROM_PY = SRC == 1, GAUSS = SRC == 2;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long ta = gw * kEmRun;
  if (ta >= E.N) return;
  const long long te = min(E.N, ta + (long long)kEmRun);
  const int U = E.U, K = E.K, Kp = E.Kp;
  bool act[GS];
  int g[GS];
  double lmu[GS][Z], l1mu[GS][Z], muC[GS][Z], cN[GS], i2v[GS], gR[GS], gX[GS];
#pragma unroll
  for (int s = 0; s < GS; ++s) {
    g[s] = lane + 32 * s;
    act[s] = g[s] < U;
    const int gg = act[s] ? g[s] : 0;
    if (!FROM_PY) {
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        lmu[s][z] = E.M.lmu[gg + z * U];
        if (!GAUSS) l1mu[s][z] = E.M.l1mu[gg + z * U];
        muC[s][z] = E.M.muC[gg + z * U];
      }
      cN[s] = E.M.cN[gg];
      i2v[s] = E.M.i2v[gg];
      if (GAUSS) { gR[s] = E.M.gR[gg]; gX[s] = E.M.gX[gg]; }
    }
  }
  // the run's per-SNP records: one coalesced load per warp, then broadcast reads from shared memory (the dependent
  // global load at the top of every iteration was the kernel's largest stall)
  __shared__ SnpRec srec[4][kEmRun];
  if (!FROM_PY) {
    if (ta + lane < te) srec[threadIdx.x >> 5][lane] = E.snp[ta + lane];
    __syncwarp();
  }
#pragma unroll 2
  for (long long t = ta; t < te; ++t) {
    double v[GS][Z];
    double mx = -CUDART_INF;
    if (FROM_PY) {
      const double *row = E.py + (size_t)t * K;
#pragma unroll
      for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int z = 0; z < Z; ++z) {
          v[s][z] = act[s] ? __ldg(row + g[s] + z * U) : -CUDART_INF;
          mx = fmax(mx, v[s][z]);
        }
    } else {
      const SnpRec r = srec[threadIdx.x >> 5][(int)(t - ta)];
#pragma unroll
      for (int s = 0; s < GS; ++s)
#pragma unroll
        for (int z = 0; z < Z; ++z) {
          const double d = r.lr - muC[s][z];
          double e;
          if (GAUSS) {
            const double dr = r.x - lmu[s][z];        // r.x = ref / tumDepth, lmu = muR
            e = cN[s] - ((dr * dr * gR[s] + d * d * i2v[s]) - dr * d * gX[s]);
          } else {
            e = (r.lb + (r.x * lmu[s][z] + r.nx * l1mu[s][z])) + (cN[s] - d * d * i2v[s]);
          }
          v[s][z] = act[s] ? e : -CUDART_INF;
          mx = fmax(mx, v[s][z]);
        }
    }
    const float mf = warp_maxf(__double2float_rn(mx));
    const double m = (mf == -CUDART_INF_F) ? 0.0 : (double)mf;     // a float-rounded shift may sit an ulp off the max: harmless
    double *dst = E.Bm + (size_t)t * Kp;
#pragma unroll
    for (int s = 0; s < GS; ++s)
      if (act[s])
#pragma unroll
        for (int z = 0; z < Z; ++z) dst[g[s] + z * U] = exp(v[s][z] - m);
    if (lane == 0) dst[K] = m;
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
bool fb2_supported(int U, int Z) { return U >= 1 && U <= 64 && Z >= 1 && Z <= 8; }

// a warp's shared memory: kNS input slots, the mbarriers, and (round kernels) two alpha tiles on their way out
static size_t ring_bytes(int TS, int Kp, bool fin) {
  size_t slot = (size_t)TS * Kp * 8 + (size_t)TS * sizeof(TxnRec);
  if (fin) slot += (size_t)TS * Kp * 8 + (size_t)TS * 32;
  size_t bytes = kNS * slot + kNS * 8 + 8;
  if (!fin) bytes += 2 * (size_t)TS * Kp * 8;
  return ((bytes + 127) / 128) * 128;
}
// Steps per tile: the largest power of two that keeps a warp's ring near 10 KB, so that six 4-warp CTAs fit an SM
// and every chunk of a 1 M-SNP genome is resident in a single wave.
static int tile_steps(int Kp, bool fin) {
  for (int ts = 8; ts > 1; ts >>= 1)
    if (ring_bytes(ts, Kp, fin) <= (fin ? 10240 : 14336)) return ts;
  return 1;
}

#define FB2_DISPATCH(Zv, GSv, ...)                                      \
  switch ((GSv) * 16 + (Zv)) {                                            \
    case 17: { constexpr int ZZ = 1, GG = 1; __VA_ARGS__; } break;               \
    case 18: { constexpr int ZZ = 2, GG = 1; __VA_ARGS__; } break;               \
    case 19: { constexpr int ZZ = 3, GG = 1; __VA_ARGS__; } break;               \
    case 20: { constexpr int ZZ = 4, GG = 1; __VA_ARGS__; } break;               \
    case 21: { constexpr int ZZ = 5, GG = 1; __VA_ARGS__; } break;               \
    case 22: { constexpr int ZZ = 6, GG = 1; __VA_ARGS__; } break;               \
    case 23: { constexpr int ZZ = 7, GG = 1; __VA_ARGS__; } break;               \
    case 24: { constexpr int ZZ = 8, GG = 1; __VA_ARGS__; } break;               \
    case 33: { constexpr int ZZ = 1, GG = 2; __VA_ARGS__; } break;               \
    case 34: { constexpr int ZZ = 2, GG = 2; __VA_ARGS__; } break;               \
    case 35: { constexpr int ZZ = 3, GG = 2; __VA_ARGS__; } break;               \
    case 36: { constexpr int ZZ = 4, GG = 2; __VA_ARGS__; } break;               \
    case 37: { constexpr int ZZ = 5, GG = 2; __VA_ARGS__; } break;               \
    case 38: { constexpr int ZZ = 6, GG = 2; __VA_ARGS__; } break;               \
    case 39: { constexpr int ZZ = 7, GG = 2; __VA_ARGS__; } break;               \
    case 40: { constexpr int ZZ = 8, GG = 2; __VA_ARGS__; } break;               \
    default: return cudaErrorInvalidValue;                                \
  }

template <class KernelT>
static cudaError_t set_smem(KernelT kern, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: set on every launch
}

cudaError_t fb2_launch_emission(const EmArgs &E, int Z, cudaStream_t st) {
  const int GS = E.U > 32 ? 2 : 1;
  const long long warps = (E.N + kEmRun - 1) / kEmRun;
  const unsigned blocks = (unsigned)((warps + 3) / 4);
  if (E.py != nullptr) { FB2_DISPATCH(Z, GS, (fb2_emission_kernel<ZZ, GG, 1><<<blocks, 128, 0, st>>>(E))); }
  else if (E.M.gauss) { FB2_DISPATCH(Z, GS, (fb2_emission_kernel<ZZ, GG, 2><<<blocks, 128, 0, st>>>(E))); }
  else { FB2_DISPATCH(Z, GS, (fb2_emission_kernel<ZZ, GG, 0><<<blocks, 128, 0, st>>>(E))); }
  return cudaGetLastError();
}

cudaError_t fb2_launch_round(const Fb2Args &A, int Z, cudaStream_t st) {
  const int GS = A.U > 32 ? 2 : 1;
  const int TS = tile_steps(A.Kp, false);
  const size_t ring = ring_bytes(TS, A.Kp, false), smem = kWPC * ring;
  const unsigned blocks = (unsigned)((2 * A.nChunks + kWPC - 1) / kWPC);
  cudaError_t e = cudaSuccess;
  FB2_DISPATCH(Z, GS, {
    e = set_smem(fb2_round_kernel<ZZ, GG>, smem);
    if (e == cudaSuccess) fb2_round_kernel<ZZ, GG><<<blocks, 32 * kWPC, smem, st>>>(A, TS, (int)ring);
  });
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t fb2_launch_probe(const Fb2Args &A, int Z, cudaStream_t st) {
  const int GS = A.U > 32 ? 2 : 1;
  const int TS = tile_steps(A.Kp, false);
  const size_t ring = ring_bytes(TS, A.Kp, false), smem = kWPC * ring;
  const unsigned blocks = (unsigned)((2 * A.nChunks * kProbes + kWPC - 1) / kWPC);
  cudaError_t e = cudaSuccess;
  FB2_DISPATCH(Z, GS, {
    e = set_smem(fb2_probe_kernel<ZZ, GG>, smem);
    if (e == cudaSuccess) fb2_probe_kernel<ZZ, GG><<<blocks, 32 * kWPC, smem, st>>>(A, TS, (int)ring);
  });
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t fb2_launch_solve(const Fb2Args &A, int Z, cudaStream_t st) {
  const int GS = A.U > 32 ? 2 : 1;
  const unsigned blocks = (unsigned)((2 * A.nChr + kWPC - 1) / kWPC);
  FB2_DISPATCH(Z, GS, (fb2_solve_kernel<ZZ, GG><<<blocks, 32 * kWPC, 0, st>>>(A)));
  return cudaGetLastError();
}

cudaError_t fb2_launch_final(const Fb2Args &A, int Z, bool write_post, bool from_py, cudaStream_t st) {
  const int GS = A.U > 32 ? 2 : 1;
  const int TS = tile_steps(A.Kp, true);
  const size_t ring = ring_bytes(TS, A.Kp, true);
  const size_t smem = kWPC * ring + (size_t)kWPC * 6 * A.K * sizeof(double);     // rings, then the quarter sums behind them
  const unsigned blocks = (unsigned)A.nChunks;
  cudaError_t e = cudaSuccess;
  if (from_py) {
    FB2_DISPATCH(Z, GS, {
      e = set_smem(fb2_final_kernel<ZZ, GG, false, true>, smem);
      if (e == cudaSuccess) fb2_final_kernel<ZZ, GG, false, true><<<blocks, 32 * kWPC, smem, st>>>(A, TS, (int)ring);
    });
  } else if (write_post) {
    FB2_DISPATCH(Z, GS, {
      e = set_smem(fb2_final_kernel<ZZ, GG, true, false>, smem);
      if (e == cudaSuccess) fb2_final_kernel<ZZ, GG, true, false><<<blocks, 32 * kWPC, smem, st>>>(A, TS, (int)ring);
    });
  } else {
    FB2_DISPATCH(Z, GS, {
      e = set_smem(fb2_final_kernel<ZZ, GG, false, false>, smem);
      if (e == cudaSuccess) fb2_final_kernel<ZZ, GG, false, false><<<blocks, 32 * kWPC, smem, st>>>(A, TS, (int)ring);
    });
  }
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace titan
