/* Minimal stand-in for the R C API, used ONLY to compile the reference's own
 * src/ C files (and this repo's R glue) in a container that has no R installation.
 * TEST INFRASTRUCTURE: nothing in the product path includes this header.
 *
 * It models an SEXP as a small tagged heap record.  Only the entry points the
 * TitanCNA C sources touch are provided.
 */
#ifndef TITAN_SHIM_R_H
#define TITAN_SHIM_R_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INTSXP  13
#define REALSXP 14
#define STRSXP  16
#define VECSXP  19
#define CHARSXP  9
#define EXTPTRSXP 22
#define LGLSXP  10
#define TRUE  1
#define FALSE 0
typedef int Rboolean;

typedef struct shim_sexp {
  int type;                 /* INTSXP / REALSXP / STRSXP / VECSXP / CHARSXP */
  int length;               /* number of elements */
  void *data;               /* double[] / int[] / SEXP[] / char[] */
  struct shim_sexp *dim;    /* INTSXP of length 2 or NULL */
  struct shim_sexp *names;  /* STRSXP or NULL */
} *SEXP;

typedef int R_len_t;
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct shim_dllinfo { const R_CallMethodDef *call; } DllInfo;

SEXP shim_alloc(int type, int n);
SEXP shim_alloc_matrix(int type, int nr, int nc);
SEXP shim_mkchar(const char *s);
void shim_error(const char *fmt, ...) __attribute__((noreturn));
int  R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);

/* external pointers + finalizers (used by the fused glue): the shim keeps a registry and runs the finalizers when
 * the test harness asks for a "garbage collection" (shim_run_finalizers) */
typedef void (*R_CFinalizer_t)(struct shim_sexp *);
struct shim_sexp *R_MakeExternalPtr(void *p, struct shim_sexp *tag, struct shim_sexp *prot);
void *R_ExternalPtrAddr(struct shim_sexp *s);
void R_ClearExternalPtr(struct shim_sexp *s);
void R_RegisterCFinalizerEx(struct shim_sexp *s, R_CFinalizer_t fun, Rboolean onexit);
int shim_run_finalizers(void);
double asReal(struct shim_sexp *x);
int asInteger(struct shim_sexp *x);
char *R_alloc(size_t n, int size);
struct shim_sexp *shim_list_get(struct shim_sexp *list, const char *name);

extern SEXP R_NamesSymbol;
extern SEXP R_NilValue;

#define PROTECT(x)            (x)
#define UNPROTECT(n)          ((void)(n))
#define AS_NUMERIC(x)         (x)
#define NUMERIC_POINTER(x)    ((double *)((x)->data))
#define REAL(x)               ((double *)((x)->data))
#define INTEGER(x)            ((int *)((x)->data))
#define INTEGER_POINTER(x)    ((int *)((x)->data))
#define GET_LENGTH(x)         ((x)->length)
#define LENGTH(x)             ((x)->length)
#define GET_DIM(x)            ((x)->dim)
#define NEW_NUMERIC(n)        shim_alloc(REALSXP, (n))
#define allocVector(t, n)     shim_alloc((t), (int)(n))
#define allocMatrix(t, r, c)  shim_alloc_matrix((t), (r), (c))
#define mkChar(s)             shim_mkchar(s)
#define SET_STRING_ELT(x,i,v) (((SEXP *)((x)->data))[i] = (v))
#define SET_VECTOR_ELT(x,i,v) (((SEXP *)((x)->data))[i] = (v))
#define VECTOR_ELT(x,i)       (((SEXP *)((x)->data))[i])
#define setAttrib(x, sym, v)  ((x)->names = (v))
#define getAttrib(x, sym)     ((x)->names ? (x)->names : R_NilValue)
#define STRING_ELT(x,i)       (((SEXP *)((x)->data))[i])
#define CHAR(x)               ((const char *)((x)->data))
#define TYPEOF(x)             ((x)->type)
#define isNull(x)             ((x) == R_NilValue || (x) == NULL)
#define error                 shim_error
#define Rf_error              shim_error

#ifdef __cplusplus
}
#endif
#endif
