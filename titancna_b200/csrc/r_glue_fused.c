/* r_glue_fused.c -- Level-2 R binding: the bodies of runEMclonalCN / viterbiClonalCN call the device-resident solver.
 *
 * Three more .Call routines next to the reference-compatible ones of r_glue.c (same shared object, same
 * R_init_TitanCNA table).  They replace the N-sized R code around the reference's per-chromosome .Call
 * (R/hmmClonal.R:141-171,191-234,261-291 for the EM; :449-484 for Viterbi): with them only O(U*Z) numbers cross
 * the boundary per EM iteration, which is what the benchmarked numbers need (the Level-1 routines still ship a
 * materialised K x N `py` per call).
 *
 *   solver <- .Call("titan_solver_create_R", chr_start, posn, ref, tumDepth, logR, txnExpLen, txnZstrength, ZS, Z)
 *   fit    <- .Call("titan_em_run_R", solver, hyper, maxiter, maxiterUpdate, normal_map, estimateS, estimatePloidy,
 *                   likChangeThreshold)              # named list of R/hmmClonal.R:344-357: n s var phi piG piZ muR muC
 *                                                    # loglik rhoG rhoZ, iteration axis trimmed to 1:i
 *   path   <- .Call("titan_viterbi_R", solver, n, s, phi, var, piG, piZ, rt, ct, rn)   # INTSXP[N], 1-based
 *
 * alleleEmissionModel = "Gaussian" (R/hmmClonal.R:149-158,265-274,439-447): titan_solver_create_gaussian_R (same
 * arguments) builds the solver for the bivariate-normal emission, titan_em_run_R then also reads hyper$varR_0,
 * alphaRHyper, betaRHyper, corRho_0, and titan_viterbi_gaussian_R takes the varR column and the correlation as two
 * more arguments.
 *
 * The solver is an external pointer whose finalizer releases the device memory when R collects it.
 * In this repository (no R installation) the file is compiled against oracle/shim/R.h and driven through the
 * ctypes .Call emulation (tests/test_gpu_parity.py::test_fused_r_glue_*).
 */
#include <R.h>
#include <Rdefines.h>
#include <Rinternals.h>
#include <stdint.h>
#include <string.h>

#include "titancna_b200.h"

static SEXP list_get(SEXP list, const char *name) {
  SEXP names = getAttrib(list, R_NamesSymbol);
  if (isNull(names)) return R_NilValue;
  for (int i = 0; i < LENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

static const double *need_vec(SEXP list, const char *name, int len) {
  SEXP v = list_get(list, name);
  if (isNull(v) || TYPEOF(v) != REALSXP || LENGTH(v) < len) error("titan_em_run_R: hyper$%s must be a numeric vector of length %d", name, len);
  return REAL(v);
}
static double need_num(SEXP list, const char *name) {
  SEXP v = list_get(list, name);
  if (isNull(v) || LENGTH(v) < 1) error("titan_em_run_R: hyper$%s is missing", name);
  return asReal(v);
}

static void solver_finalizer(SEXP ptr) {
  titan_solver *s = (titan_solver *)R_ExternalPtrAddr(ptr);
  if (s) {
    titan_solver_destroy(s);
    R_ClearExternalPtr(ptr);
  }
}

static titan_solver *solver_of(SEXP ptr, const char *who) {
  titan_solver *s = (TYPEOF(ptr) == EXTPTRSXP) ? (titan_solver *)R_ExternalPtrAddr(ptr) : NULL;
  if (!s) error("%s: not a live titan solver (was it created by titan_solver_create_R?)", who);
  return s;
}

/* chr_start: offsets [nChr+1] of the contiguous chromosome runs of data$chr (0-based, last = N) */
static SEXP solver_create(SEXP chr_start, SEXP posn, SEXP ref, SEXP depth, SEXP logR, SEXP txnExpLen, SEXP txnZstrength,
                          SEXP ZS, SEXP Z, int alleleModel) {
  PROTECT(chr_start = AS_NUMERIC(chr_start));
  PROTECT(posn = AS_NUMERIC(posn));
  PROTECT(ref = AS_NUMERIC(ref));
  PROTECT(depth = AS_NUMERIC(depth));
  PROTECT(logR = AS_NUMERIC(logR));
  PROTECT(ZS = AS_NUMERIC(ZS));
  const int nchr = LENGTH(chr_start) - 1, N = LENGTH(posn);
  if (nchr < 1 || LENGTH(ref) != N || LENGTH(depth) != N || LENGTH(logR) != N)
    error("titan_solver_create_R: data must contain named list elements: chr, posn, ref, tumDepth, logR of equal length");
  int64_t *cs = (int64_t *)R_alloc((size_t)nchr + 1, sizeof(int64_t));
  for (int i = 0; i <= nchr; ++i) cs[i] = (int64_t)REAL(chr_start)[i];
  titan_data_desc d;
  memset(&d, 0, sizeof d);
  d.N = N; d.nChr = nchr; d.chr_start = cs;
  d.posn = REAL(posn); d.ref = REAL(ref); d.tumDepth = REAL(depth); d.logR = REAL(logR);
  d.txnExpLen = asReal(txnExpLen); d.txnZstrength = asReal(txnZstrength);
  d.U = LENGTH(ZS); d.Z = asInteger(Z); d.ZS = REAL(ZS);
  d.alleleModel = alleleModel;
  titan_solver *s = NULL;
  if (titan_solver_create(&d, &s) != TITAN_OK) error("titan_solver_create: %s", titan_last_error());
  SEXP ptr = PROTECT(R_MakeExternalPtr(s, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, solver_finalizer, TRUE);
  UNPROTECT(7);
  return ptr;
}

SEXP titan_solver_create_R(SEXP chr_start, SEXP posn, SEXP ref, SEXP depth, SEXP logR, SEXP txnExpLen, SEXP txnZstrength,
                           SEXP ZS, SEXP Z) {
  return solver_create(chr_start, posn, ref, depth, logR, txnExpLen, txnZstrength, ZS, Z, TITAN_ALLELE_BINOMIAL);
}
SEXP titan_solver_create_gaussian_R(SEXP chr_start, SEXP posn, SEXP ref, SEXP depth, SEXP logR, SEXP txnExpLen,
                                    SEXP txnZstrength, SEXP ZS, SEXP Z) {
  return solver_create(chr_start, posn, ref, depth, logR, txnExpLen, txnZstrength, ZS, Z, TITAN_ALLELE_GAUSSIAN);
}

static SEXP named_real(SEXP list, SEXP names, int slot, const char *name, int nr, int nc) {
  SEXP v = (nc > 0) ? allocMatrix(REALSXP, nr, nc) : allocVector(REALSXP, nr);
  SET_VECTOR_ELT(list, slot, v);                 /* reachable from the protected list from here on */
  SET_STRING_ELT(names, slot, mkChar(name));
  return v;
}

/* hyper: one flat named list with the fields of loadDefaultParameters() the M-step and the priors read
 * (c(genotypeParams, normalParams, ploidyParams, cellPrevParams) in R): rt ct ZS rn var_0 alphaKHyper betaKHyper
 * kappaGHyper piG_0 s_0 alphaSHyper betaSHyper kappaZHyper piZ_0 n_0 alphaNHyper betaNHyper phi_0 alphaPHyper betaPHyper. */
SEXP titan_em_run_R(SEXP solver, SEXP hyper, SEXP maxiter, SEXP maxiterUpdate, SEXP normal_map, SEXP estimateS,
                    SEXP estimatePloidy, SEXP likChangeThreshold) {
  titan_solver *s = solver_of(solver, "titan_em_run_R");
  SEXP rt = list_get(hyper, "rt"), s0 = list_get(hyper, "s_0");
  if (isNull(rt) || isNull(s0)) error("titan_em_run_R: params must include genotypeParams, ploidyParams, normalParams, cellPrevParams.");
  const int U = LENGTH(rt), Z = LENGTH(s0), K = U * Z, M = asInteger(maxiter) + 1;
  int64_t N64 = 0;
  int sU = 0, sZ = 0;
  if (titan_solver_shape(s, &N64, NULL, &sU, &sZ) != TITAN_OK) error("titan_em_run_R: %s", titan_last_error());
  if (sU != U || sZ != Z) error("titan_em_run_R: the solver was created for %d genotype states and %d clusters, params hold %d and %d", sU, sZ, U, Z);
  const int N = (int)N64;
  titan_hyper h;
  memset(&h, 0, sizeof h);
  h.rt = need_vec(hyper, "rt", U); h.ct = need_vec(hyper, "ct", U); h.ZS = need_vec(hyper, "ZS", U);
  h.rn = need_num(hyper, "rn");
  h.var_0 = need_vec(hyper, "var_0", U); h.alphaKHyper = need_vec(hyper, "alphaKHyper", U);
  h.betaKHyper = need_vec(hyper, "betaKHyper", U); h.kappaGHyper = need_vec(hyper, "kappaGHyper", U);
  h.piG_0 = need_vec(hyper, "piG_0", U);
  h.s_0 = need_vec(hyper, "s_0", Z); h.alphaSHyper = need_vec(hyper, "alphaSHyper", Z); h.betaSHyper = need_vec(hyper, "betaSHyper", Z);
  h.kappaZHyper = need_vec(hyper, "kappaZHyper", Z); h.piZ_0 = need_vec(hyper, "piZ_0", Z);
  h.n_0 = need_num(hyper, "n_0"); h.alphaNHyper = need_num(hyper, "alphaNHyper"); h.betaNHyper = need_num(hyper, "betaNHyper");
  h.phi_0 = need_num(hyper, "phi_0"); h.alphaPHyper = need_num(hyper, "alphaPHyper"); h.betaPHyper = need_num(hyper, "betaPHyper");
  /* optional: read by a Gaussian solver (which refuses to run without them), carried along by a binomial one */
  if (!isNull(list_get(hyper, "varR_0"))) h.varR_0 = need_vec(hyper, "varR_0", U);
  if (!isNull(list_get(hyper, "alphaRHyper"))) h.alphaRHyper = need_vec(hyper, "alphaRHyper", U);
  if (!isNull(list_get(hyper, "betaRHyper"))) h.betaRHyper = need_vec(hyper, "betaRHyper", U);
  if (!isNull(list_get(hyper, "corRho_0"))) h.corRho_0 = need_num(hyper, "corRho_0");
  titan_em_opts o;
  memset(&o, 0, sizeof o);
  o.maxiter = asInteger(maxiter); o.maxiterUpdate = asInteger(maxiterUpdate); o.normal_map = asInteger(normal_map);
  o.estimateS = asInteger(estimateS); o.estimatePloidy = asInteger(estimatePloidy);
  o.likChangeThreshold = asReal(likChangeThreshold); o.want_posteriors = 1; o.verbose = 0;

  /* scratch with the full iteration axis (R_alloc: released by R at the end of the .Call) */
  titan_em_out out;
  memset(&out, 0, sizeof out);
  out.n = (double *)R_alloc(M, sizeof(double)); out.phi = (double *)R_alloc(M, sizeof(double));
  out.loglik = (double *)R_alloc(M, sizeof(double));
  out.s = (double *)R_alloc((size_t)Z * M, sizeof(double)); out.piZ = (double *)R_alloc((size_t)Z * M, sizeof(double));
  out.var = (double *)R_alloc((size_t)U * M, sizeof(double)); out.piG = (double *)R_alloc((size_t)U * M, sizeof(double));
  out.muR = (double *)R_alloc((size_t)K * M, sizeof(double)); out.muC = (double *)R_alloc((size_t)K * M, sizeof(double));
  out.varR = (double *)R_alloc((size_t)U * M, sizeof(double));

  enum { F_n, F_s, F_var, F_varR, F_phi, F_piG, F_piZ, F_muR, F_muC, F_loglik, F_rhoG, F_rhoZ, F_count };
  SEXP list = PROTECT(allocVector(VECSXP, F_count)), names = PROTECT(allocVector(STRSXP, F_count));
  SEXP rhoG = named_real(list, names, F_rhoG, "rhoG", U, N), rhoZ = named_real(list, names, F_rhoZ, "rhoZ", Z, N);
  out.rhoG = REAL(rhoG);
  out.rhoZ = REAL(rhoZ);
  if (titan_em_run(s, &h, &o, &out) != TITAN_OK) error("titan_em_run: %s", titan_last_error());
  const int I = out.iters;            /* the R code's final i: columns 1:i are valid (R/hmmClonal.R:344-353) */
  memcpy(REAL(named_real(list, names, F_n, "n", I, 0)), out.n, sizeof(double) * I);
  memcpy(REAL(named_real(list, names, F_phi, "phi", I, 0)), out.phi, sizeof(double) * I);
  memcpy(REAL(named_real(list, names, F_loglik, "loglik", I, 0)), out.loglik, sizeof(double) * I);
  memcpy(REAL(named_real(list, names, F_s, "s", Z, I)), out.s, sizeof(double) * (size_t)Z * I);
  memcpy(REAL(named_real(list, names, F_piZ, "piZ", Z, I)), out.piZ, sizeof(double) * (size_t)Z * I);
  memcpy(REAL(named_real(list, names, F_var, "var", U, I)), out.var, sizeof(double) * (size_t)U * I);
  memcpy(REAL(named_real(list, names, F_varR, "varR", U, I)), out.varR, sizeof(double) * (size_t)U * I);
  memcpy(REAL(named_real(list, names, F_piG, "piG", U, I)), out.piG, sizeof(double) * (size_t)U * I);
  /* muR / muC: [U, Z, i] arrays; handed over as (U*Z) x i matrices, R sets dim <- c(U, Z, i) */
  memcpy(REAL(named_real(list, names, F_muR, "muR", K, I)), out.muR, sizeof(double) * (size_t)K * I);
  memcpy(REAL(named_real(list, names, F_muC, "muC", K, I)), out.muC, sizeof(double) * (size_t)K * I);
  setAttrib(list, R_NamesSymbol, names);
  UNPROTECT(2);
  return list;
}

/* Emission parameters from column i, initial distribution from column i-1 (R/hmmClonal.R:402,466-467): the caller
 * passes exactly those vectors. */
static SEXP viterbi_call(SEXP solver, SEXP n, SEXP s_, SEXP phi, SEXP var, SEXP piG, SEXP piZ, SEXP rt, SEXP ct, SEXP rn,
                         SEXP varR, SEXP corRho) {
  titan_solver *s = solver_of(solver, "titan_viterbi_R");
  PROTECT(s_ = AS_NUMERIC(s_));
  PROTECT(var = AS_NUMERIC(var));
  PROTECT(piG = AS_NUMERIC(piG));
  PROTECT(piZ = AS_NUMERIC(piZ));
  PROTECT(rt = AS_NUMERIC(rt));
  PROTECT(ct = AS_NUMERIC(ct));
  const int U = LENGTH(rt), Z = LENGTH(s_);
  if (LENGTH(var) != U || LENGTH(piG) != U || LENGTH(ct) != U || LENGTH(piZ) != Z)
    error("titan_viterbi_R: var, piG, ct must have length(rt) entries and piZ length(s)");
  titan_params p;
  memset(&p, 0, sizeof p);
  p.n = asReal(n); p.phi = asReal(phi); p.s = REAL(s_); p.var = REAL(var); p.piG = REAL(piG); p.piZ = REAL(piZ);
  p.rt = REAL(rt); p.ct = REAL(ct); p.rn = asReal(rn);
  if (varR != R_NilValue) {
    PROTECT(varR = AS_NUMERIC(varR));
    if (LENGTH(varR) != U) error("titan_viterbi_gaussian_R: varR must have length(rt) entries");
    p.varR = REAL(varR);
    p.corRho = asReal(corRho);
  }
  int64_t N = 0;
  if (titan_solver_shape(s, &N, NULL, NULL, NULL) != TITAN_OK) error("titan_viterbi_R: %s", titan_last_error());
  SEXP path = PROTECT(allocVector(INTSXP, (R_len_t)N));
  if (titan_viterbi(s, &p, (int32_t *)INTEGER(path)) != TITAN_OK) error("titan_viterbi: %s", titan_last_error());
  UNPROTECT(varR != R_NilValue ? 8 : 7);
  return path;
}

SEXP titan_viterbi_R(SEXP solver, SEXP n, SEXP s_, SEXP phi, SEXP var, SEXP piG, SEXP piZ, SEXP rt, SEXP ct, SEXP rn) {
  return viterbi_call(solver, n, s_, phi, var, piG, piZ, rt, ct, rn, R_NilValue, R_NilValue);
}
/* corRho: the reference's viterbiClonalCN re-derives it from the data (R/hmmClonal.R:447,536-538) */
SEXP titan_viterbi_gaussian_R(SEXP solver, SEXP n, SEXP s_, SEXP phi, SEXP var, SEXP piG, SEXP piZ, SEXP rt, SEXP ct, SEXP rn,
                              SEXP varR, SEXP corRho) {
  return viterbi_call(solver, n, s_, phi, var, piG, piZ, rt, ct, rn, varR, corRho);
}
