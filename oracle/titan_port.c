/* titan_port.c -- plain-C CPU restatement of TitanCNA's native hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: tests/, the smoke test and
 * bench.py's cpu_baseline leg call it; nothing under titancna_b200/ links or loads it.
 *
 * It restates, in this repo's own words, the dense log-space algorithm of
 *   /root/reference/src/fwd_backC_clonalCN.c   (forward-backward, :40-211; transition
 *                                               matrix :215-301; distance term :342-348;
 *                                               logsumexp :440-457)
 *   /root/reference/src/viterbiC_clonalCN.c    (Viterbi :35-134; arg-max :210-229)
 * keeping the reference's floating-point evaluation order wherever the order is
 * observable (sequential row sums, max-then-sum logsumexp, strict-greater arg-max,
 * single additions in delta), so that Viterbi paths can be compared bit for bit.
 * Parity of this restatement against the reference's own object code is pinned in
 * tests/test_oracle_pinning.py (oracle/_ref/libtitanref.so, built from the reference
 * sources by oracle/Makefile) and against the reference fixture data/EMresults.rda
 * (tests/golden/).
 *
 * The only liberty taken is a one-entry cache of the logged transition matrix keyed on
 * the exact (rhoG, rhoZ) bit patterns: consecutive SNP gaps often repeat, and rebuilding
 * an identical matrix returns identical bits.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* reference: distanceTransitionFunction, fwd_backC_clonalCN.c:342-348 */
double port_rho(double prev, double cur, double L) {
  double gap = cur - prev + 1.0;
  return (1.0 / 2.0) * (1.0 - exp(-gap / (2.0 * L)));
}

/* Zygosity class / cluster number of joint state index s.
 * reference: fwd_backC_clonalCN.c:224-247 (outlier variant :225-232) */
static void state_class(int s, int U, const double *ZS, int outlier, double *zs, double *clu) {
  if (outlier) {
    if (s == 0) { *zs = -100.0; *clu = 0.0; return; }
    *clu = ceil(((double)s) / U);
    *zs = ZS[(s - 1) % U];
  } else {
    *clu = ceil(((double)s + 1) / U);
    *zs = ZS[s % U];
  }
}

/* Row-normalised K x K transition matrix, column-major with rows = source state:
 * A[i + j*K] = P(i -> j).  reference: preparePositionSpecificMatrix :215-301 */
void port_transition(double *A, int K, int U, const double *ZS, double rhoG, double rhoZ, int outlier) {
  double offG = (1.0 - rhoG) / ((double)K - 1.0);
  double offZ = 1.0 - rhoZ;
  for (int i = 0; i < K; ++i) {
    double zsi, ci;
    state_class(i, U, ZS, outlier, &zsi, &ci);
    for (int j = 0; j < K; ++j) {
      double zsj, cj;
      state_class(j, U, ZS, outlier, &zsj, &cj);
      double g = (zsi == zsj) ? rhoG : offG;
      A[i + (size_t)j * K] = (ci == cj || zsj == -1) ? g * rhoZ : g * offZ;
    }
  }
  for (int i = 0; i < K; ++i) {
    double tot = 0;
    for (int j = 0; j < K; ++j) tot += A[i + (size_t)j * K];   /* sequential, in j order */
    if (tot > 0)
      for (int j = 0; j < K; ++j) A[i + (size_t)j * K] /= tot;
  }
}

/* reference: logsumexp :440-457 -- strict '>' max scan, then sequential sum of exp */
static double lse(const double *x, int n) {
  double mx = x[0], acc = 0.0;
  for (int i = 0; i < n; ++i) if (x[i] > mx) mx = x[i];
  for (int i = 0; i < n; ++i) acc += exp(x[i] - mx);
  return mx + log(acc);
}

static double lse_normalise(double *x, int n) {
  double z = lse(x, n);
  for (int i = 0; i < n; ++i) x[i] -= z;
  return z;
}

typedef struct {
  int K, U, outlier, valid;
  const double *ZS;
  double kG, kZ;
  double *logA;   /* log P(i->j) at [i + j*K] */
} txn_cache;

static const double *logged_transition(txn_cache *c, double rhoG, double rhoZ) {
  if (c->valid && memcmp(&c->kG, &rhoG, sizeof rhoG) == 0 && memcmp(&c->kZ, &rhoZ, sizeof rhoZ) == 0)
    return c->logA;
  port_transition(c->logA, c->K, c->U, c->ZS, rhoG, rhoZ, c->outlier);
  size_t n = (size_t)c->K * c->K;
  for (size_t q = 0; q < n; ++q) c->logA[q] = log(c->logA[q]);
  c->kG = rhoG; c->kZ = rhoZ; c->valid = 1;
  return c->logA;
}

/* Dense log-space forward-backward for one chromosome.
 * logpi[K], py[K x T] column-major (state fastest), posn[T].
 * Writes log posteriors gamma[K x T] and the log-likelihood.  Returns 0, or -1 on bad args
 * / allocation failure.  reference: fwd_backC_clonalCN.c:111-192 */
int port_fwd_back(const double *logpi, const double *py, int K, int T, const double *ZS, int Z,
                  const double *posn, double zStrength, double txnLen, int outlier,
                  double *gamma, double *loglik) {
  if (K <= 0 || T <= 0 || Z <= 0) return -1;
  int U = K / Z;                                    /* :71 integer division */
  size_t KT = (size_t)K * T;
  double *alpha = (double *)malloc(KT * sizeof(double));
  double *beta = (double *)malloc(2 * (size_t)K * sizeof(double));  /* only t and t+1 are live */
  double *scale = (double *)malloc((size_t)T * sizeof(double));
  double *tmp = (double *)malloc((size_t)K * sizeof(double));
  double *msg = (double *)malloc((size_t)K * sizeof(double));
  txn_cache c = {K, U, outlier, 0, ZS, 0, 0, (double *)malloc((size_t)K * K * sizeof(double))};
  if (!alpha || !beta || !scale || !tmp || !msg || !c.logA) {
    free(alpha); free(beta); free(scale); free(tmp); free(msg); free(c.logA);
    return -1;
  }

  for (int k = 0; k < K; ++k) alpha[k] = logpi[k] + py[k];
  scale[0] = lse_normalise(alpha, K);
  for (int t = 1; t < T; ++t) {
    double rhoG = 1.0 - port_rho(posn[t - 1], posn[t], txnLen);
    double rhoZ = 1.0 - port_rho(posn[t - 1], posn[t], zStrength);
    const double *LA = logged_transition(&c, rhoG, rhoZ);
    const double *prev = alpha + (size_t)(t - 1) * K;
    double *cur = alpha + (size_t)t * K;
    for (int d = 0; d < K; ++d) {                   /* destination d, sum over sources i */
      for (int i = 0; i < K; ++i) tmp[i] = LA[i + (size_t)d * K] + prev[i];
      msg[d] = lse(tmp, K);
    }
    for (int k = 0; k < K; ++k) cur[k] = msg[k] + py[(size_t)t * K + k];
    scale[t] = lse_normalise(cur, K);
  }
  double ll = 0;
  for (int t = 0; t < T; ++t) ll += scale[t];
  *loglik = ll;

  double *bnext = beta, *bcur = beta + K;
  for (int k = 0; k < K; ++k) { bnext[k] = 0; gamma[(size_t)(T - 1) * K + k] = alpha[(size_t)(T - 1) * K + k]; }
  for (int t = T - 2; t >= 0; --t) {
    double rhoG = 1.0 - port_rho(posn[t], posn[t + 1], txnLen);
    double rhoZ = 1.0 - port_rho(posn[t], posn[t + 1], zStrength);
    const double *LA = logged_transition(&c, rhoG, rhoZ);
    for (int k = 0; k < K; ++k) msg[k] = bnext[k] + py[(size_t)(t + 1) * K + k];
    for (int d = 0; d < K; ++d) {                   /* source d, sum over destinations i */
      for (int i = 0; i < K; ++i) tmp[i] = LA[d + (size_t)i * K] + msg[i];
      bcur[d] = lse(tmp, K);
    }
    lse_normalise(bcur, K);
    double *g = gamma + (size_t)t * K;
    for (int k = 0; k < K; ++k) g[k] = alpha[(size_t)t * K + k] + bcur[k];
    lse_normalise(g, K);
    double *sw = bnext; bnext = bcur; bcur = sw;
  }
  free(alpha); free(beta); free(scale); free(tmp); free(msg); free(c.logA);
  return 0;
}

/* Log-space Viterbi for one chromosome; path is 1-based like the reference's return.
 * reference: viterbiC_clonalCN.c:90-128, arg-max :210-229 (strict '<', first index wins) */
int port_viterbi(const double *logpi, const double *py, int K, int T, const double *ZS, int Z,
                 const double *posn, double zStrength, double txnLen, int outlier, int *path) {
  if (K <= 0 || T <= 0 || Z <= 0) return -1;
  int U = K / Z;
  double *dprev = (double *)malloc((size_t)K * sizeof(double));
  double *dcur = (double *)malloc((size_t)K * sizeof(double));
  int *psi = (int *)malloc((size_t)K * T * sizeof(int));
  txn_cache c = {K, U, outlier, 0, ZS, 0, 0, (double *)malloc((size_t)K * K * sizeof(double))};
  if (!dprev || !dcur || !psi || !c.logA) { free(dprev); free(dcur); free(psi); free(c.logA); return -1; }

  for (int k = 0; k < K; ++k) { dprev[k] = logpi[k] + py[k]; psi[k] = 0; }
  for (int t = 1; t < T; ++t) {
    double rhoG = 1.0 - port_rho(posn[t - 1], posn[t], txnLen);
    double rhoZ = 1.0 - port_rho(posn[t - 1], posn[t], zStrength);
    const double *LA = logged_transition(&c, rhoG, rhoZ);
    for (int j = 0; j < K; ++j) {
      const double *col = LA + (size_t)j * K;
      double best = dprev[0] + col[0];
      int arg = 0;
      for (int i = 1; i < K; ++i) {
        double cand = dprev[i] + col[i];
        if (best < cand) { best = cand; arg = i; }
      }
      dcur[j] = best + py[(size_t)t * K + j];
      psi[(size_t)t * K + j] = arg;
    }
    double *sw = dprev; dprev = dcur; dcur = sw;
  }
  int arg = 0;
  for (int i = 1; i < K; ++i) if (dprev[arg] < dprev[i]) arg = i;
  path[T - 1] = arg;
  for (int t = T - 2; t >= 0; --t) path[t] = psi[(size_t)(t + 1) * K + path[t + 1]];
  for (int t = 0; t < T; ++t) path[t] += 1;
  free(dprev); free(dcur); free(psi); free(c.logA);
  return 0;
}

/* out[n] = lgamma(n + 1) for n = 0..nmax with the host libm (the table both the oracle and
 * the product's host side index; reference: binomialpdflog R/hmmClonal.R:617-619) */
void port_lgamma_table(int nmax, double *out) {
  for (int n = 0; n <= nmax; ++n) out[n] = lgamma((double)n + 1.0);
}

/* K x N log-emission matrix in the reference's association order.
 * reference: R/hmmClonal.R:491-523 (loops), :617-634 (densities), :171 (+pseudoCounts).
 * lbin[t] = (lgamma(N+1) - lgamma(x+1)) - lgamma(N-x+1) precomputed by the caller. */
void port_emission(const double *x, const double *depth, const double *logR, const double *lbin,
                   int N, const double *muR, const double *muC, const double *var, int U, int Z,
                   double pseudo, double *py) {
  int K = U * Z;
  double *lmu = (double *)malloc(sizeof(double) * K * 2);
  double *l1mu = lmu + K;
  double *cN = (double *)malloc(sizeof(double) * U * 2);
  double *tv = cN + U;
  const double two_pi = 2 * 3.141592653589793;   /* R's pi */
  for (int j = 0; j < K; ++j) { lmu[j] = log(muR[j]); l1mu[j] = log(1 - muR[j]); }
  for (int u = 0; u < U; ++u) { cN[u] = log(1 / (sqrt(var[u]) * sqrt(two_pi))); tv[u] = 2 * var[u]; }
  for (int t = 0; t < N; ++t) {
    double k = x[t], nk = depth[t] - x[t], l = logR[t];
    for (int j = 0; j < K; ++j) {
      int u = j % U;
      double pyR = lbin[t] + (k * lmu[j] + nk * l1mu[j]);
      double dlt = l - muC[j];
      double pyC = cN[u] + (-(dlt * dlt) / tv[u]);
      py[(size_t)t * K + j] = (pyR + pyC) + pseudo;
    }
  }
  free(lmu); free(cN);
}
