/* r_glue.c -- the R side of the drop-in boundary.
 *
 * Compiled by `R CMD SHLIB` (with libtitancna_b200.so on the link line) this file IS TitanCNA.so:
 * it exports R_init_TitanCNA and registers the same three .Call routines with the same arities
 * as the reference's src/register.c:8-18, so `useDynLib(TitanCNA)` (NAMESPACE:1),
 * `library.dynam` (R/zzz.R:1-3) and `.Call("fwd_backC_clonalCN", ...)` by string
 * (R/hmmClonal.R:209,478) keep working without touching the R code.  Each routine unpacks its
 * SEXPs exactly as the reference does (src/fwd_backC_clonalCN.c:48-71, src/viterbiC_clonalCN.c:45-67),
 * calls the plain C ABI of include/titancna_b200.h and repacks the result into the reference's
 * return shape (list(rho=K x T REALSXP, loglik=REALSXP[1]); INTSXP[T]).
 *
 * In this repository (no R installation) the same file is compiled against oracle/shim/R.h and
 * driven through ctypes, which is how tests/ exercise the .Call contract.
 */
#include <R.h>
#include <Rdefines.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "titancna_b200.h"

SEXP fwd_backC_clonalCN(SEXP piGiZi, SEXP py, SEXP copyNumKey, SEXP zygosityKey, SEXP numClust, SEXP positions,
                        SEXP zStrength, SEXP txnLen, SEXP useOutlier) {
  PROTECT(piGiZi = AS_NUMERIC(piGiZi));
  PROTECT(py = AS_NUMERIC(py));
  PROTECT(copyNumKey = AS_NUMERIC(copyNumKey));
  PROTECT(zygosityKey = AS_NUMERIC(zygosityKey));
  PROTECT(numClust = AS_NUMERIC(numClust));
  PROTECT(positions = AS_NUMERIC(positions));
  PROTECT(zStrength = AS_NUMERIC(zStrength));
  PROTECT(txnLen = AS_NUMERIC(txnLen));
  PROTECT(useOutlier = AS_NUMERIC(useOutlier));
  const int K = GET_LENGTH(piGiZi), T = GET_LENGTH(positions);
  if (GET_DIM(py) == NULL || INTEGER(GET_DIM(py))[0] != K || INTEGER(GET_DIM(py))[1] != T)
    error("fwd_backC_clonalCN: The obslik must be %d-by-%d dimension.", K, T);
  /* length-0 scalars are undefined behaviour in the reference: reject them */
  if (GET_LENGTH(numClust) < 1 || GET_LENGTH(zStrength) < 1 || GET_LENGTH(txnLen) < 1 || GET_LENGTH(useOutlier) < 1)
    error("fwd_backC_clonalCN: numClust, zStrength, txnLen and useOutlier must be scalars.");
  SEXP gamma_data, loglik_data, list, list_names;
  PROTECT(gamma_data = allocMatrix(REALSXP, K, T));
  PROTECT(loglik_data = NEW_NUMERIC(1));
  int rc = titan_fwd_back_clonalCN(NUMERIC_POINTER(piGiZi), K, NUMERIC_POINTER(py), T, NUMERIC_POINTER(copyNumKey),
                                   GET_LENGTH(copyNumKey), NUMERIC_POINTER(zygosityKey), GET_LENGTH(zygosityKey),
                                   (int)NUMERIC_POINTER(numClust)[0], NUMERIC_POINTER(positions),
                                   NUMERIC_POINTER(zStrength)[0], NUMERIC_POINTER(txnLen)[0],
                                   (int)NUMERIC_POINTER(useOutlier)[0], NUMERIC_POINTER(gamma_data),
                                   NUMERIC_POINTER(loglik_data));
  if (rc != TITAN_OK) error("fwd_backC_clonalCN: %s", titan_last_error());
  const char *names[2] = {"rho", "loglik"};
  PROTECT(list_names = allocVector(STRSXP, 2));
  for (int i = 0; i < 2; ++i) SET_STRING_ELT(list_names, i, mkChar(names[i]));
  PROTECT(list = allocVector(VECSXP, 2));
  SET_VECTOR_ELT(list, 0, gamma_data);
  SET_VECTOR_ELT(list, 1, loglik_data);
  setAttrib(list, R_NamesSymbol, list_names);
  UNPROTECT(13);
  return list;
}

SEXP viterbiC_clonalCN(SEXP piGiZi, SEXP py, SEXP copyNumKey, SEXP zygosityKey, SEXP numClust, SEXP positions,
                       SEXP zStrength, SEXP txnLen, SEXP useOutlier) {
  PROTECT(piGiZi = AS_NUMERIC(piGiZi));
  PROTECT(py = AS_NUMERIC(py));
  PROTECT(copyNumKey = AS_NUMERIC(copyNumKey));
  PROTECT(zygosityKey = AS_NUMERIC(zygosityKey));
  PROTECT(numClust = AS_NUMERIC(numClust));
  PROTECT(positions = AS_NUMERIC(positions));
  PROTECT(zStrength = AS_NUMERIC(zStrength));
  PROTECT(txnLen = AS_NUMERIC(txnLen));
  PROTECT(useOutlier = AS_NUMERIC(useOutlier));
  const int K = GET_LENGTH(piGiZi), T = GET_LENGTH(positions);
  if (GET_DIM(py) == NULL || INTEGER(GET_DIM(py))[0] != K || INTEGER(GET_DIM(py))[1] != T)
    error("viterbiC_clonalCN: The obslik must be %d-by-%d dimension.", K, T);
  if (GET_LENGTH(numClust) < 1 || GET_LENGTH(zStrength) < 1 || GET_LENGTH(txnLen) < 1 || GET_LENGTH(useOutlier) < 1)
    error("viterbiC_clonalCN: numClust, zStrength, txnLen and useOutlier must be scalars.");
  SEXP path_data;
  PROTECT(path_data = allocVector(INTSXP, T));
  int rc = titan_viterbi_clonalCN(NUMERIC_POINTER(piGiZi), K, NUMERIC_POINTER(py), T, NUMERIC_POINTER(copyNumKey),
                                  GET_LENGTH(copyNumKey), NUMERIC_POINTER(zygosityKey), GET_LENGTH(zygosityKey),
                                  (int)NUMERIC_POINTER(numClust)[0], NUMERIC_POINTER(positions),
                                  NUMERIC_POINTER(zStrength)[0], NUMERIC_POINTER(txnLen)[0],
                                  (int)NUMERIC_POINTER(useOutlier)[0], INTEGER_POINTER(path_data));
  if (rc != TITAN_OK) error("viterbiC_clonalCN: %s", titan_last_error());
  UNPROTECT(10);
  return path_data;
}

SEXP getPositionOverlapC(SEXP posn, SEXP start, SEXP stop) {
  PROTECT(posn = AS_NUMERIC(posn));
  PROTECT(start = AS_NUMERIC(start));
  PROTECT(stop = AS_NUMERIC(stop));
  const int T = GET_LENGTH(posn), Tcn = GET_LENGTH(start);
  SEXP out;
  PROTECT(out = allocVector(REALSXP, T));
  if (titan_get_position_overlap(NUMERIC_POINTER(posn), T, NUMERIC_POINTER(start), NUMERIC_POINTER(stop), Tcn,
                                 NUMERIC_POINTER(out)) != TITAN_OK)
    error("getPositionOverlapC: %s", titan_last_error());
  UNPROTECT(4);
  return out;
}

/* r_glue_fused.c: the solver-based routines the patched bodies of runEMclonalCN / viterbiClonalCN call */
SEXP titan_solver_create_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP titan_em_run_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP titan_viterbi_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP titan_solver_create_gaussian_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP titan_viterbi_gaussian_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* the reference's table (src/register.c:8-13) first, unchanged; then the solver routines */
static const R_CallMethodDef callMethods[] = {
    {"getPositionOverlapC", (DL_FUNC)&getPositionOverlapC, 3},
    {"fwd_backC_clonalCN", (DL_FUNC)&fwd_backC_clonalCN, 9},
    {"viterbiC_clonalCN", (DL_FUNC)&viterbiC_clonalCN, 9},
    {"titan_solver_create_R", (DL_FUNC)&titan_solver_create_R, 9},
    {"titan_em_run_R", (DL_FUNC)&titan_em_run_R, 8},
    {"titan_viterbi_R", (DL_FUNC)&titan_viterbi_R, 10},
    {"titan_solver_create_gaussian_R", (DL_FUNC)&titan_solver_create_gaussian_R, 9},
    {"titan_viterbi_gaussian_R", (DL_FUNC)&titan_viterbi_gaussian_R, 12},
    {NULL, NULL, 0}};

void R_init_TitanCNA(DllInfo *info) { R_registerRoutines(info, NULL, callMethods, NULL, NULL); }
