// titan_host.hpp -- host-side (CPU, O(U*Z) or parameter-independent O(N)) parts of the path.
//
// Everything here is either
//   (a) parameter-independent preprocessing that MUST use the host libm to reproduce the
//       reference bit for bit (rhoG/rhoZ at txnExpLen = 1e15 are quantised to a few multiples
//       of 2^-53; Viterbi log-transition tables), or
//   (b) the O(U*Z) scalar M-step of R/paramEstimation.R, which the reference also runs on the
//       host (interpreted R) and which is not N-sized work.
// None of it is a fallback for the device kernels.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace titan_host {

static const double kPi = 3.141592653589793;   // R's `pi`

// reference: distanceTransitionFunction, src/fwd_backC_clonalCN.c:342-348; callers :124-125
inline double rho_keep(double prev, double cur, double L) {
  double distance = cur - prev + 1.0;
  double rho = (1.0 / 2.0) * (1.0 - std::exp(-distance / (2.0 * L)));
  return 1.0 - rho;
}

// reference: clonalTwoComponentMixtureCN, R/hmmClonal.R:579-603 (same association order)
inline void mixture_means(int U, int Z, const double *rt, double rn, double n, const double *s, const double *ct,
                          double phi, double *muR, double *muC) {
  const double cn = 2.0;
  for (int k = 0; k < U; ++k)
    for (int z = 0; z < Z; ++z) {
      double numRef = n * rn * cn + (1 - n) * s[z] * rn * cn + (1 - n) * (1 - s[z]) * rt[k] * ct[k];
      double total = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * ct[k];
      double ploidy = n * cn + (1 - n) * phi;
      muR[k + z * U] = numRef / total;
      muC[k + z * U] = std::log(total / ploidy);
    }
}

// lgamma(n+1) table at integer arguments (binomialpdflog's normalising constant, R/hmmClonal.R:618)
struct LgammaTable {
  std::vector<double> v;
  void ensure(int64_t nmax) {
    int64_t have = (int64_t)v.size();
    if (have > nmax) return;
    v.resize((size_t)nmax + 1);
    for (int64_t i = have; i <= nmax; ++i) v[(size_t)i] = std::lgamma((double)i + 1.0);
  }
};

inline double binom_log_norm(double x, double depth, LgammaTable &tab) {
  double xi = std::floor(x), di = std::floor(depth);
  if (xi == x && di == depth && x >= 0 && depth >= x && depth < 1e7) {
    tab.ensure((int64_t)depth);
    return (tab.v[(size_t)depth] - tab.v[(size_t)x]) - tab.v[(size_t)(depth - x)];
  }
  return (std::lgamma(depth + 1) - std::lgamma(x + 1)) - std::lgamma(depth - x + 1);
}

inline uint64_t bits_of(double d);
// ---- Viterbi: exact log-transition table for one (rhoG, rhoZ) pair ---------------------------
// out[i*4 + cls] = log(raw_cls / rowsum_i), cls = 2*sameGenotype + (sameCluster || destHET).
// rowsum_i is accumulated sequentially over j exactly as preparePositionSpecificMatrix does
// (src/fwd_backC_clonalCN.c:283-292), the log is the host libm's (logMatrixInPlace :460-465).
inline void viterbi_log_table(int K, int U, const unsigned char *het, double rhoG, double rhoZ, double *out) {
  const double gdiff = (1.0 - rhoG) / ((double)K - 1.0);
  double raw[4];
  raw[0] = gdiff * (1.0 - rhoZ);
  raw[1] = gdiff * rhoZ;
  raw[2] = rhoG * (1.0 - rhoZ);
  raw[3] = rhoG * rhoZ;
  for (int i = 0; i < K; ++i) {
    const int gi = i % U, zi = i / U;
    double sum = 0;
    int gj = 0, zj = 0;
    for (int j = 0; j < K; ++j) {
      int cls = ((gi == gj) ? 2 : 0) | ((zi == zj || het[gj]) ? 1 : 0);
      sum += raw[cls];
      if (++gj == U) { gj = 0; ++zj; }
    }
    for (int c = 0; c < 4; ++c) out[i * 4 + c] = std::log(sum > 0 ? raw[c] / sum : raw[c]);
  }
}

// The same table for many pairs: the class of every (i, j) is computed once (structured_class_matrix), the row sums
// of eight rows advance together -- each still adds its K terms in the reference's order j = 0 .. K-1, so every sum
// is bit-identical to viterbi_log_table's, but the eight dependent chains overlap instead of one add waiting for the
// previous one -- and the four libm logs of a row are reused when the next row's sum has the same bits.
inline void structured_class_matrix(int K, int U, const unsigned char *het, std::vector<unsigned char> &cls) {
  cls.assign((size_t)K * K, 0);
  for (int i = 0; i < K; ++i) {
    const int gi = i % U, zi = i / U;
    for (int j = 0; j < K; ++j) {
      const int gj = j % U, zj = j / U;
      cls[(size_t)i * K + j] = (unsigned char)(((gi == gj) ? 2 : 0) | ((zi == zj || het[gj]) ? 1 : 0));
    }
  }
}

inline void viterbi_log_table_rows8(int K, const unsigned char *cls, double rhoG, double rhoZ, double *out) {
  const double gdiff = (1.0 - rhoG) / ((double)K - 1.0);
  const double raw[4] = {gdiff * (1.0 - rhoZ), gdiff * rhoZ, rhoG * (1.0 - rhoZ), rhoG * rhoZ};
  double prev_sum = -1.0, prev_log[4] = {0, 0, 0, 0};
  for (int i0 = 0; i0 < K; i0 += 8) {
    const int nr = K - i0 < 8 ? K - i0 : 8;
    double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned char *row[8];
    for (int r = 0; r < 8; ++r) row[r] = cls + (size_t)(i0 + (r < nr ? r : 0)) * K;
    for (int j = 0; j < K; ++j) {
      sum[0] += raw[row[0][j]]; sum[1] += raw[row[1][j]]; sum[2] += raw[row[2][j]]; sum[3] += raw[row[3][j]];
      sum[4] += raw[row[4][j]]; sum[5] += raw[row[5][j]]; sum[6] += raw[row[6][j]]; sum[7] += raw[row[7][j]];
    }
    for (int r = 0; r < nr; ++r) {
      double *o = out + (size_t)(i0 + r) * 4;
      if (bits_of(sum[r]) != bits_of(prev_sum)) {
        for (int c = 0; c < 4; ++c) prev_log[c] = std::log(sum[r] > 0 ? raw[c] / sum[r] : raw[c]);
        prev_sum = sum[r];
      }
      o[0] = prev_log[0]; o[1] = prev_log[1]; o[2] = prev_log[2]; o[3] = prev_log[3];
    }
  }
}

// Zygosity class / cluster number of joint state s exactly as the reference derives them
// (src/fwd_backC_clonalCN.c:224-247), and the class of every transition i -> j:
// 2*(zygosity_i == zygosity_j) + (cluster_i == cluster_j || zygosity_j == -1).
inline void generic_class_matrix(int K, int U, const double *ZS, bool outlier, std::vector<unsigned char> &cls) {
  std::vector<double> zs(K), clu(K);
  for (int s = 0; s < K; ++s) {
    if (outlier) {
      if (s == 0) { zs[s] = -100.0; clu[s] = 0.0; }
      else { clu[s] = std::ceil(((double)s) / U); zs[s] = ZS[(s - 1) % U]; }
    } else {
      clu[s] = std::ceil(((double)s + 1) / U);
      zs[s] = ZS[s % U];
    }
  }
  cls.assign((size_t)K * K, 0);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      cls[(size_t)i * K + j] = (unsigned char)(((zs[i] == zs[j]) ? 2 : 0) | ((clu[i] == clu[j] || zs[j] == -1) ? 1 : 0));
}

// viterbi_log_table for an arbitrary class matrix (same arithmetic: sequential row sums, libm log)
inline void viterbi_log_table_generic(int K, const unsigned char *cls, double rhoG, double rhoZ, double *out) {
  const double gdiff = (1.0 - rhoG) / ((double)K - 1.0);
  const double raw[4] = {gdiff * (1.0 - rhoZ), gdiff * rhoZ, rhoG * (1.0 - rhoZ), rhoG * rhoZ};
  for (int i = 0; i < K; ++i) {
    double sum = 0;
    for (int j = 0; j < K; ++j) sum += raw[cls[(size_t)i * K + j]];
    for (int c = 0; c < 4; ++c) out[i * 4 + c] = std::log(sum > 0 ? raw[c] / sum : raw[c]);
  }
}

struct PairHash {
  size_t operator()(const std::pair<uint64_t, uint64_t> &p) const {
    return std::hash<uint64_t>()(p.first * 0x9E3779B97F4A7C15ull ^ (p.second + (p.first << 6) + (p.first >> 2)));
  }
};

inline uint64_t bits_of(double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return u;
}

// ---- M-step -----------------------------------------------------------------------------------
struct Hyper {
  int U = 0, Z = 0;
  std::vector<double> rt, ct, var_0, alphaK, betaK, kappaG, piG_0, s_0, alphaS, betaS, kappaZ, piZ_0;
  double rn = 0.5, n_0 = 0.5, alphaN = 2, betaN = 2, phi_0 = 2, alphaP = 20, betaP = 42;
  // alleleEmissionModel == "Gaussian" (R/utils.R:108-119): allelic variance, its prior, the fixed correlation
  bool gauss = false;
  std::vector<double> varR_0, alphaR, betaR;
  double corRho = 0;
};

struct Stats {   // each [K], joint index u + z*U  (R/paramEstimation.R:34-40)
  const double *a, *b, *c, *d, *e, *g;
};

inline double lbeta_(double a, double b) { return std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b); }
// R/hmmClonal.R:646-657
inline double betapdflog(double x, double a, double b) {
  return -lbeta_(a, b) + (a - 1) * std::log(x) + (b - 1) * std::log(1 - x);
}
inline double invgammapdflog(double x, double a, double b) {
  return (a * std::log(b) - std::lgamma(a)) + ((-a - 1) * std::log(x) + (-b / x));
}
// R/hmmClonal.R:659-665
inline double dirichletpdflog(const double *x, const double *k, int n) {
  double sk = 0, sl = 0, l = 0;
  for (int i = 0; i < n; ++i) { sk += k[i]; sl += std::lgamma(k[i]); }
  for (int i = 0; i < n; ++i) l += (k[i] - 1) * std::log(x[i]);
  return (std::lgamma(sk) - sl) + l;
}

struct Priors { double prior, s, n, phi, var, piG, piZ, varR; };

// R/hmmClonal.R:715-783; the varR term (:767-775) only for the Gaussian allele model
inline Priors prior_probs(const Hyper &h, double n, const double *s, double phi, const double *var, const double *piG,
                          const double *piZ, bool normal_map, bool estS, bool estP, const double *varR = nullptr) {
  Priors p{};
  p.piG = dirichletpdflog(piG, h.kappaG.data(), h.U);
  p.piZ = dirichletpdflog(piZ, h.kappaZ.data(), h.Z);
  if (estS)
    for (int z = 0; z < h.Z; ++z) p.s += betapdflog(s[z], h.alphaS[z], h.betaS[z]);
  p.n = normal_map ? betapdflog(n, h.alphaN, h.betaN) : 0.0;
  for (int k = 0; k < h.U; ++k) {          // first state of every copy-number level (:746-756)
    bool first = true;
    for (int q = 0; q < k; ++q)
      if (h.ct[q] == h.ct[k]) { first = false; break; }
    if (first) p.var += invgammapdflog(var[k], h.alphaK[k], h.betaK[k]);
  }
  p.phi = estP ? invgammapdflog(phi, h.alphaP, h.betaP) : 0.0;
  if (h.gauss && varR)
    for (int k = 0; k < h.U; ++k) p.varR += invgammapdflog(varR[k], h.alphaR[k], h.betaR[k]);
  p.prior = p.s + p.n + p.phi + p.var + p.varR + p.piG + p.piZ;
  return p;
}

// R/paramEstimation.R:287-356, binomial branch
inline double likelihood_func(const Hyper &h, const Stats &st, double n, const double *s, const double *var, double phi,
                              const double *piG, const double *piZ, double likInitG, double likInitZ, bool normal_map,
                              bool estS, bool estP) {
  const int U = h.U, Z = h.Z;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  double lik = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const int j = k + z * U;
      lik = lik + st.d[j] + st.a[j] * std::log(muR[j]) + st.b[j] * std::log(1 - muR[j]) +
            st.e[j] * std::log(1 / std::sqrt(2 * kPi * var[k])) -
            (st.g[j] - 2 * st.c[j] * muC[j] + st.e[j] * (muC[j] * muC[j])) / (2 * var[k]);
    }
  Priors p = prior_probs(h, n, s, phi, var, piG, piZ, normal_map, estS, estP);
  return lik + p.prior + likInitG + likInitZ + 0.0;
}

// R/paramEstimation.R:359-399 (binomial branch)
inline double deriv_n(const Hyper &h, const Stats &st, double n, const double *s, const double *var, double phi) {
  const int U = h.U, Z = h.Z;
  const double cn = 2.0;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  double tot = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const int j = k + z * U;
      double den = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * h.ct[k];
      double dmuC = (cn - s[z] * cn - (1 - s[z]) * h.ct[k]) / den - (cn - phi) / (n * cn + (1 - n) * phi);
      double dmuR = ((1 - s[z]) * cn * h.ct[k] * (h.rn - h.rt[k])) / (den * den);
      double t1 = st.a[j] * dmuR / muR[j] - st.b[j] * dmuR / (1 - muR[j]);
      double t2 = (st.c[j] - st.e[j] * muC[j]) * dmuC / var[k];
      tot = tot + (t1 + t2);
    }
  return tot + ((h.alphaN - 1) / n - (h.betaN - 1) / (1 - n));
}

// R/paramEstimation.R:401-437 for cluster z (binomial branch)
inline double deriv_s(const Hyper &h, const Stats &st, double sz, int z, double n, const double *var, double phi) {
  const int U = h.U;
  const double cn = 2.0;
  double tot = 0;
  for (int k = 0; k < U; ++k) {
    const int j = k + z * U;
    double numRef = n * h.rn * cn + (1 - n) * sz * h.rn * cn + (1 - n) * (1 - sz) * h.rt[k] * h.ct[k];
    double den = n * cn + (1 - n) * sz * cn + (1 - n) * (1 - sz) * h.ct[k];
    double muR = numRef / den;
    double muC = std::log(den / (n * cn + (1 - n) * phi));
    double dmuC = ((1 - n) * cn - (1 - n) * h.ct[k]) / den;
    double dmuR = ((1 - n) * cn * h.ct[k] * (h.rn - h.rt[k])) / (den * den);
    double t1 = st.a[j] * dmuR / muR - st.b[j] * dmuR / (1 - muR);
    double t2 = (st.c[j] - st.e[j] * muC) * dmuC / var[k];
    tot = tot + (t1 + t2);
  }
  return tot + ((h.alphaS[z] - 1) / sz - (h.betaS[z] - 1) / (1 - sz));
}

// R/paramEstimation.R:442-461
inline double var_exact(const Hyper &h, const Stats &st, double n, const double *s, double phi, double ck, double alphaK,
                        double betaK) {
  const int U = h.U, Z = h.Z;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  double t1 = 0, t2 = 0;
  for (int k = 0; k < U; ++k) {
    if (h.ct[k] != ck) continue;
    for (int z = 0; z < Z; ++z) {
      const int j = k + z * U;
      t1 = t1 + -(st.g[j] - 2 * st.c[j] * muC[j] + st.e[j] * (muC[j] * muC[j]));
      t2 = t2 + -(st.e[j]);
    }
  }
  t1 = t1 - 2 * betaK;
  t2 = t2 - 2 * (alphaK + 1);
  return t1 / t2;
}

// R/paramEstimation.R:499-530 (binomial branch)
inline double deriv_phi(const Hyper &h, const Stats &st, double phi, double n, const double *s, const double *var) {
  const int U = h.U, Z = h.Z;
  const double cn = 2.0;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  const double dmuC = -(1 - n) / (n * cn + (1 - n) * phi);
  double tot = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const int j = k + z * U;
      double t1 = st.c[j] / var[k];
      double t2 = st.e[j] * muC[j] / var[k];
      tot = tot + (t1 - t2) * dmuC;
    }
  return tot + ((-h.alphaP - 1) / phi + h.betaP / (phi * phi));
}

// ---- Gaussian allele model: the other branch of the same R functions ---------------------------
// Statistics (R/paramEstimation.R:29-33,37-39): a = sum (x/N) rho, f = sum (x/N)^2 rho, h = sum (x/N) l rho, c, e, g.
// The reference's own slips are kept as written: log(1 / 2 * pi * ...) in the objective (:321) and the single
// subscript a[k], c[k], e[k], mus$R[k], mus$C[k] (= column 1 for every cluster) in the ploidy derivative (:512-513).
struct GStats {
  const double *a, *c, *e, *f, *g, *h;   // each [K], joint index u + z*U
};

// R/paramEstimation.R:287-356, Gaussian branch :313-329
inline double likelihood_func_gauss(const Hyper &h, const GStats &st, double n, const double *s, const double *var,
                                    const double *varR, double phi, const double *piG, const double *piZ, double likInitG,
                                    double likInitZ, bool normal_map, bool estS, bool estP) {
  const int U = h.U, Z = h.Z;
  const double rho = h.corRho;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  double lik = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const int j = k + z * U;
      const double term1 = st.e[j] * std::log(((1.0 / 2) * kPi) * std::sqrt((var[k] * varR[k]) * (1 - rho * rho)));
      const double term2 = 1 / (2 * (1 - rho * rho));
      const double term3 = ((st.f[j] - (2 * st.a[j]) * muR[j]) + st.e[j] * (muR[j] * muR[j])) / (2 * varR[k]);
      const double term4 = ((st.g[j] - (2 * st.c[j]) * muC[j]) + st.e[j] * (muC[j] * muC[j])) / (2 * var[k]);
      const double term5 = ((2 * rho) * (((st.h[j] - st.a[j] * muC[j]) - st.c[j] * muR[j]) + (st.e[j] * muC[j]) * muR[j])) /
                           std::sqrt(var[k] * varR[k]);
      lik = (lik + term1) - term2 * ((term3 + term4) - term5);
    }
  Priors p = prior_probs(h, n, s, phi, var, piG, piZ, normal_map, estS, estP, varR);
  return lik + p.prior + likInitG + likInitZ + 0.0;
}

// R/paramEstimation.R:359-399, Gaussian branch :376-383
inline double deriv_n_gauss(const Hyper &h, const GStats &st, double n, const double *s, const double *var,
                            const double *varR, double phi) {
  const int U = h.U, Z = h.Z;
  const double cn = 2.0, rho = h.corRho;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  double tot = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const int j = k + z * U;
      double den = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * h.ct[k];
      double dmuC = (cn - s[z] * cn - (1 - s[z]) * h.ct[k]) / den - (cn - phi) / (n * cn + (1 - n) * phi);
      double dmuR = ((1 - s[z]) * cn * h.ct[k] * (h.rn - h.rt[k])) / (den * den);
      const double term0 = 1 / (1 - rho * rho);
      const double rR = st.a[j] - st.e[j] * muR[j], rC = st.c[j] - st.e[j] * muC[j];
      const double term1 = rR * dmuR / varR[k];
      const double term2 = rC * dmuC / var[k];
      const double term3 = rho * ((rC * dmuR) + (rR * dmuC)) / std::sqrt(varR[k] * var[k]);
      tot = tot + term0 * (term1 + term2 - term3);
    }
  return tot + ((h.alphaN - 1) / n - (h.betaN - 1) / (1 - n));
}

// R/paramEstimation.R:401-437 for cluster z, Gaussian branch :415-422
inline double deriv_s_gauss(const Hyper &h, const GStats &st, double sz, int z, double n, const double *var,
                            const double *varR, double phi) {
  const int U = h.U;
  const double cn = 2.0, rho = h.corRho;
  double tot = 0;
  for (int k = 0; k < U; ++k) {
    const int j = k + z * U;
    double numRef = n * h.rn * cn + (1 - n) * sz * h.rn * cn + (1 - n) * (1 - sz) * h.rt[k] * h.ct[k];
    double den = n * cn + (1 - n) * sz * cn + (1 - n) * (1 - sz) * h.ct[k];
    double muR = numRef / den;
    double muC = std::log(den / (n * cn + (1 - n) * phi));
    double dmuC = ((1 - n) * cn - (1 - n) * h.ct[k]) / den;
    double dmuR = ((1 - n) * cn * h.ct[k] * (h.rn - h.rt[k])) / (den * den);
    const double term0 = 1 / (1 - rho * rho);
    const double rR = st.a[j] - st.e[j] * muR, rC = st.c[j] - st.e[j] * muC;
    const double term1 = rR * dmuR / varR[k];
    const double term2 = rC * dmuC / var[k];
    const double term3 = rho * ((rC * dmuR) + (rR * dmuC)) / std::sqrt(varR[k] * var[k]);
    tot = tot + term0 * (term1 + term2 - term3);
  }
  return tot + ((h.alphaS[z] - 1) / sz - (h.betaS[z] - 1) / (1 - sz));
}

// clonalCNDerivativeVarUpdateEqn, R/paramEstimation.R:463-497, over the genotype states ind[0..nInd).
// logRatio (:168-176): unknown = var of a copy-number level, var2 = varR_prev of the same states.
// allelicRatio (:187-192): unknown = varR[k] (nInd = 1), var2 = the NEW var[k]; the caller's swap of a <-> c and
// g := f is done here.  The inverse-gamma prior derivative is summed over all nInd states (:493-494) with R's
// long-double sum().
inline double deriv_var_gauss(const Hyper &h, const GStats &st, double v, double n, const double *s, const double *var2,
                              double phi, const int *ind, int nInd, bool logratio) {
  const int U = h.U, Z = h.Z;
  const double cn = 2.0, rho = h.corRho;
  double dlik = 0;
  for (int z = 0; z < Z; ++z)
    for (int q = 0; q < nInd; ++q) {
      const int k = ind[q], j = k + z * U;
      double numRef = n * h.rn * cn + (1 - n) * s[z] * h.rn * cn + (1 - n) * (1 - s[z]) * h.rt[k] * h.ct[k];
      double den = n * cn + (1 - n) * s[z] * cn + (1 - n) * (1 - s[z]) * h.ct[k];
      const double muR = numRef / den, muC = std::log(den / (n * cn + (1 - n) * phi));
      const double mus1 = logratio ? muC : muR, mus2 = logratio ? muR : muC;
      const double A = logratio ? st.a[j] : st.c[j], Cc = logratio ? st.c[j] : st.a[j];
      const double G = logratio ? st.g[j] : st.f[j];
      const double term1 = st.e[j] / (2 * v);
      const double term2 = 1 / (2 * (1 - rho * rho));
      const double term3 = ((G - (2 * Cc) * mus1) + st.e[j] * (mus1 * mus1)) / (v * v);
      const double term4 = (rho * (((st.h[j] - Cc * mus2) - A * mus1) + (st.e[j] * mus1) * mus2)) / (v * std::sqrt(v * var2[q]));
      dlik = (dlik + (-term1)) + (term2 * (term3 - term4));
    }
  long double dinv = 0;
  for (int q = 0; q < nInd; ++q) {
    const int k = ind[q];
    const double aK = logratio ? h.alphaK[k] : h.alphaR[k], bK = logratio ? h.betaK[k] : h.betaR[k];
    dinv += (long double)((-aK - 1) / v + bK / (v * v));
  }
  return dlik + (double)dinv;
}

// R/paramEstimation.R:499-530, Gaussian branch :510-515 (single-subscript indexing = cluster 1, repeated Z times)
inline double deriv_phi_gauss(const Hyper &h, const GStats &st, double phi, double n, const double *s, const double *var,
                              const double *varR) {
  const int U = h.U, Z = h.Z;
  const double cn = 2.0, rho = h.corRho;
  std::vector<double> muR((size_t)U * Z), muC((size_t)U * Z);
  mixture_means(U, Z, h.rt.data(), h.rn, n, s, h.ct.data(), phi, muR.data(), muC.data());
  const double dmuC = -(1 - n) / (n * cn + (1 - n) * phi);
  double tot = 0;
  for (int z = 0; z < Z; ++z)
    for (int k = 0; k < U; ++k) {
      const double term0 = 1 / (1 - rho * rho);
      const double term1 = (st.c[k] - st.e[k] * muC[k]) * dmuC / var[k];
      const double term2 = rho * ((st.a[k] - st.e[k] * muR[k]) * dmuC) / std::sqrt(varR[k] * var[k]);
      tot = tot + term0 * (term1 - term2);
    }
  return tot + ((-h.alphaP - 1) / phi + h.betaP / (phi * phi));
}

// Brent's zeroin as used by stats::uniroot (R core zeroin.c, R_zeroin2); R core is not vendored in
// the reference checkout, so this follows the published Forsythe-Malcolm-Moler / netlib form.
template <class F>
inline double r_zeroin(F &&f, double ax, double bx, double fa, double fb, double tol, int maxit) {
  const double EPS = std::numeric_limits<double>::epsilon();
  double a = ax, b = bx, c = ax, fc = fa;
  if (fa == 0.0) return a;
  if (fb == 0.0) return b;
  int it = maxit + 1;
  while (it-- > 0) {
    double prev_step = b - a;
    if (std::fabs(fc) < std::fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    double tol_act = 2 * EPS * std::fabs(b) + tol / 2;
    double new_step = (c - b) / 2;
    if (std::fabs(new_step) <= tol_act || fb == 0.0) return b;
    if (std::fabs(prev_step) >= tol_act && std::fabs(fa) > std::fabs(fb)) {
      double p, q, t1, t2, cb = c - b;
      if (a == c) {
        t1 = fb / fa; p = cb * t1; q = 1.0 - t1;
      } else {
        q = fa / fc; t1 = fb / fc; t2 = fb / fa;
        p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
        q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
      }
      if (p > 0) q = -q; else p = -p;
      if (p < (0.75 * cb * q - std::fabs(tol_act * q) / 2) && p < std::fabs(prev_step * q / 2)) new_step = p / q;
    }
    if (std::fabs(new_step) < tol_act) new_step = (new_step > 0) ? tol_act : -tol_act;
    a = b; fa = fb;
    b += new_step;
    fb = f(b);
    if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) { c = a; fc = fa; }
  }
  return b;
}

// stats::uniroot front end; returns false where uniroot would stop() (the reference's tryCatch
// then keeps the previous iterate, R/paramEstimation.R:134-139,154-159,222-227)
template <class F>
inline bool uniroot(F &&f, double lower, double upper, double *root) {
  const double big = std::numeric_limits<double>::max();
  bool nan_seen = false;
  auto g = [&](double x) {
    double v = f(x);
    if (std::isnan(v)) nan_seen = true;
    return std::fmax(std::fmin(v, big), -big);
  };
  double fl = f(lower), fu = f(upper);
  if (std::isnan(fl) || std::isnan(fu)) return false;
  fl = std::fmax(std::fmin(fl, big), -big);
  fu = std::fmax(std::fmin(fu, big), -big);
  auto sgn = [](double v) { return (v > 0) - (v < 0); };
  if (!(sgn(fl) * sgn(fu) <= 0)) return false;
  double r = r_zeroin(g, lower, upper, fl, fu, 1e-15, 1000);
  if (nan_seen) return false;
  *root = r;
  return true;
}

struct MStepOut {
  double n, phi;
  std::vector<double> s, var, varR;
  int sweeps;
  double maxF;
};

// Coordinate descent of R/paramEstimation.R:75-283 (binomial model): per sweep n -> s_1..s_Z ->
// var per copy-number level -> phi, loop `while ((!converged && i < maxiter) || i <= miniter)`
// with miniter = 10, and the LAST sweep is returned (:272).
inline MStepOut update_parameters(const Hyper &h, const Stats &st, double n_0, const double *s_0, const double *var_0,
                                  double phi_0, const double *piG, const double *piZ, double likInitG, double likInitZ,
                                  bool normal_map, bool estS, bool estP, int maxiter, int miniter = 10) {
  const int U = h.U, Z = h.Z;
  const double EPS = std::numeric_limits<double>::epsilon();
  double n_prev = n_0, phi_prev = phi_0;
  std::vector<double> s_prev(s_0, s_0 + Z), var_prev(var_0, var_0 + U);
  auto F = [&](double n, const double *s, const double *var, double phi) {
    return likelihood_func(h, st, n, s, var, phi, piG, piZ, likInitG, likInitZ, normal_map, estS, estP);
  };
  double obj_prev = F(n_prev, s_prev.data(), var_prev.data(), phi_prev), obj = obj_prev;
  bool converged = false;
  int i = 1;
  double n = n_prev, phi = phi_prev;
  std::vector<double> s = s_prev, var = var_prev;
  while ((!converged && i < maxiter) || i <= miniter) {
    ++i;
    if (normal_map) {
      double r;
      auto fn = [&](double v) { return deriv_n(h, st, v, s_prev.data(), var_prev.data(), phi_prev); };
      n = uniroot(fn, EPS, 1 - EPS, &r) ? r : n_prev;
    } else {
      n = n_prev;
    }
    s = s_prev;
    if (estS)
      for (int z = 0; z < Z; ++z) {
        double r;
        auto fs = [&](double v) { return deriv_s(h, st, v, z, n, var_prev.data(), phi_prev); };
        s[z] = uniroot(fs, EPS, 1 - EPS, &r) ? r : s_prev[z];
      }
    std::vector<double> varn(U, 0.0);
    for (int k = 0; k < U; ++k) {
      bool first = true;
      for (int q = 0; q < k; ++q)
        if (h.ct[q] == h.ct[k]) { first = false; varn[k] = varn[q]; break; }
      if (first) varn[k] = var_exact(h, st, n, s.data(), phi_prev, h.ct[k], h.alphaK[k], h.betaK[k]);
    }
    var = varn;
    if (estP) {
      double r;
      auto fp = [&](double v) { return deriv_phi(h, st, v, n, s.data(), var.data()); };
      phi = uniroot(fp, EPS, 10.0, &r) ? r : phi_prev;
    } else {
      phi = phi_prev;
    }
    obj = F(n, s.data(), var.data(), phi);
    n_prev = n; s_prev = s; var_prev = var; phi_prev = phi;
    if (!std::isnan(obj) && (std::fabs(obj - obj_prev) / std::fabs(obj)) <= 0.001) converged = true;
    obj_prev = obj;
  }
  MStepOut o;
  o.n = n; o.phi = phi; o.s = s; o.var = var; o.sweeps = i - 1; o.maxF = obj;
  return o;
}

// Coordinate descent of R/paramEstimation.R:75-283, Gaussian allele model: per sweep n -> s_1..s_Z -> var per
// copy-number level (uniroot on (eps, 5), :165-183) -> varR per genotype state (:185-200, with the NEW var[k]) -> phi.
inline MStepOut update_parameters_gauss(const Hyper &h, const GStats &st, double n_0, const double *s_0, const double *var_0,
                                        const double *varR_0, double phi_0, const double *piG, const double *piZ,
                                        double likInitG, double likInitZ, bool normal_map, bool estS, bool estP, int maxiter,
                                        int miniter = 10) {
  const int U = h.U, Z = h.Z;
  const double EPS = std::numeric_limits<double>::epsilon();
  double n_prev = n_0, phi_prev = phi_0;
  std::vector<double> s_prev(s_0, s_0 + Z), var_prev(var_0, var_0 + U), varR_prev(varR_0, varR_0 + U);
  auto F = [&](double n, const double *s, const double *var, const double *varR, double phi) {
    return likelihood_func_gauss(h, st, n, s, var, varR, phi, piG, piZ, likInitG, likInitZ, normal_map, estS, estP);
  };
  double obj_prev = F(n_prev, s_prev.data(), var_prev.data(), varR_prev.data(), phi_prev), obj = obj_prev;
  bool converged = false;
  int i = 1;
  double n = n_prev, phi = phi_prev;
  std::vector<double> s = s_prev, var = var_prev, varR = varR_prev;
  while ((!converged && i < maxiter) || i <= miniter) {
    ++i;
    if (normal_map) {
      double r;
      auto fn = [&](double v) { return deriv_n_gauss(h, st, v, s_prev.data(), var_prev.data(), varR_prev.data(), phi_prev); };
      n = uniroot(fn, EPS, 1 - EPS, &r) ? r : n_prev;
    } else {
      n = n_prev;
    }
    s = s_prev;
    if (estS)
      for (int z = 0; z < Z; ++z) {
        double r;
        auto fs = [&](double v) { return deriv_s_gauss(h, st, v, z, n, var_prev.data(), varR_prev.data(), phi_prev); };
        s[z] = uniroot(fs, EPS, 1 - EPS, &r) ? r : s_prev[z];
      }
    std::vector<double> varn(U, 0.0);
    std::vector<char> done(U, 0);
    for (int k = 0; k < U; ++k) {
      if (done[k]) continue;
      std::vector<int> ind;
      std::vector<double> v2;
      for (int q = k; q < U; ++q)
        if (h.ct[q] == h.ct[k]) { ind.push_back(q); v2.push_back(varR_prev[q]); done[q] = 1; }
      double r;
      auto fv = [&](double v) { return deriv_var_gauss(h, st, v, n, s.data(), v2.data(), phi_prev, ind.data(), (int)ind.size(), true); };
      const bool ok = uniroot(fv, EPS, 5.0, &r);
      for (int q : ind) varn[q] = ok ? r : var_prev[q];
    }
    var = varn;
    for (int k = 0; k < U; ++k) {
      double r;
      const double v2 = var[k];
      auto fv = [&](double v) { return deriv_var_gauss(h, st, v, n, s.data(), &v2, phi_prev, &k, 1, false); };
      varR[k] = uniroot(fv, EPS, 5.0, &r) ? r : varR_prev[k];
    }
    if (estP) {
      double r;
      auto fp = [&](double v) { return deriv_phi_gauss(h, st, v, n, s.data(), var.data(), varR.data()); };
      phi = uniroot(fp, EPS, 10.0, &r) ? r : phi_prev;
    } else {
      phi = phi_prev;
    }
    obj = F(n, s.data(), var.data(), varR.data(), phi);
    n_prev = n; s_prev = s; var_prev = var; varR_prev = varR; phi_prev = phi;
    if (!std::isnan(obj) && (std::fabs(obj - obj_prev) / std::fabs(obj)) <= 0.001) converged = true;
    obj_prev = obj;
  }
  MStepOut o;
  o.n = n; o.phi = phi; o.s = s; o.var = var; o.varR = varR; o.sweeps = i - 1; o.maxF = obj;
  return o;
}

}  // namespace titan_host
