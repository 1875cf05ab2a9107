/* Runtime for the shim R.h: arena allocation, error() via longjmp, and
 * guarded call trampolines so a ctypes caller survives error().
 * TEST INFRASTRUCTURE ONLY. */
#include "R.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct shim_sexp names_symbol, nil_value;
SEXP R_NamesSymbol = &names_symbol;
SEXP R_NilValue = &nil_value;

/* per-thread state: bench.py drives one chromosome per worker thread (foreach %dopar% analogue) */
static __thread void **arena = NULL;
static __thread size_t arena_n = 0, arena_cap = 0;
static __thread jmp_buf shim_jmp;
static __thread int shim_jmp_armed = 0;
static __thread char shim_msg[1024];
static const R_CallMethodDef *shim_registered = NULL;

static void *arena_calloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "shim: out of memory\n"); abort(); }
  if (arena_n == arena_cap) {
    arena_cap = arena_cap ? 2 * arena_cap : 256;
    arena = (void **)realloc(arena, arena_cap * sizeof(void *));
  }
  arena[arena_n++] = p;
  return p;
}

void shim_free_all(void) {
  for (size_t i = 0; i < arena_n; ++i) free(arena[i]);
  arena_n = 0;
}

static size_t elt_size(int type) {
  switch (type) {
    case INTSXP: return sizeof(int);
    case REALSXP: return sizeof(double);
    case CHARSXP: return 1;
    default: return sizeof(SEXP);
  }
}

SEXP shim_alloc(int type, int n) {
  SEXP s = (SEXP)arena_calloc(1, sizeof(*s));
  s->type = type; s->length = n;
  s->data = arena_calloc((size_t)n, elt_size(type));
  return s;
}

SEXP shim_alloc_matrix(int type, int nr, int nc) {
  SEXP s = shim_alloc(type, nr * nc);
  s->dim = shim_alloc(INTSXP, 2);
  INTEGER(s->dim)[0] = nr; INTEGER(s->dim)[1] = nc;
  return s;
}

/* wrap caller-owned memory without copying (inputs from numpy) */
SEXP shim_wrap(int type, int n, void *data, int nr, int nc) {
  SEXP s = (SEXP)arena_calloc(1, sizeof(*s));
  s->type = type; s->length = n; s->data = data;
  if (nr >= 0) {
    s->dim = shim_alloc(INTSXP, 2);
    INTEGER(s->dim)[0] = nr; INTEGER(s->dim)[1] = nc;
  }
  return s;
}

SEXP shim_mkchar(const char *str) {
  int n = (int)strlen(str);
  SEXP s = shim_alloc(CHARSXP, n + 1);
  memcpy(s->data, str, (size_t)n + 1);
  return s;
}

void shim_error(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(shim_msg, sizeof shim_msg, fmt, ap);
  va_end(ap);
  if (shim_jmp_armed) longjmp(shim_jmp, 1);
  fprintf(stderr, "shim error(): %s\n", shim_msg);
  abort();
}

const char *shim_last_error(void) { return shim_msg; }

/* ---- external pointers: heap records that survive shim_free_all (R would keep them until garbage collection) ---- */
#define SHIM_MAX_EXTPTR 256
static struct { SEXP s; R_CFinalizer_t fun; } extptrs[SHIM_MAX_EXTPTR];
static int n_extptrs = 0;

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  s->type = EXTPTRSXP; s->length = 1; s->data = p;
  return s;
}
void *R_ExternalPtrAddr(SEXP s) { return (s && s->type == EXTPTRSXP) ? s->data : NULL; }
void R_ClearExternalPtr(SEXP s) { if (s) s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
  (void)onexit;
  if (n_extptrs < SHIM_MAX_EXTPTR) { extptrs[n_extptrs].s = s; extptrs[n_extptrs].fun = fun; ++n_extptrs; }
}
/* "gc()": run every registered finalizer once, free the records; returns how many ran */
int shim_run_finalizers(void) {
  int n = n_extptrs;
  for (int i = 0; i < n; ++i) { extptrs[i].fun(extptrs[i].s); free(extptrs[i].s); }
  n_extptrs = 0;
  return n;
}
double asReal(SEXP x) { return x->type == INTSXP || x->type == LGLSXP ? (double)((int *)x->data)[0] : ((double *)x->data)[0]; }
int asInteger(SEXP x) { return x->type == INTSXP || x->type == LGLSXP ? ((int *)x->data)[0] : (int)((double *)x->data)[0]; }
char *R_alloc(size_t n, int size) { return (char *)arena_calloc(n, (size_t)size); }
/* list element by name (the usual getListElement helper of R extensions); R_NilValue when absent */
SEXP shim_list_get(SEXP list, const char *name) {
  if (!list || list->type != VECSXP || !list->names) return R_NilValue;
  for (int i = 0; i < list->length; ++i)
    if (strcmp((const char *)((SEXP *)list->names->data)[i]->data, name) == 0) return ((SEXP *)list->data)[i];
  return R_NilValue;
}
/* build a named list from the harness side */
SEXP shim_named_list(int n, const char **names, SEXP *elts) {
  SEXP l = shim_alloc(VECSXP, n), nm = shim_alloc(STRSXP, n);
  for (int i = 0; i < n; ++i) { ((SEXP *)l->data)[i] = elts[i]; ((SEXP *)nm->data)[i] = shim_mkchar(names[i]); }
  l->names = nm;
  return l;
}

int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call,
                       const void *f, const void *e) {
  (void)c; (void)f; (void)e;
  if (info) info->call = call;
  shim_registered = call;
  return 1;
}

/* look a routine up in the table R_init_<pkg> registered; -1 arity if absent */
DL_FUNC shim_lookup(const char *name, int *nargs) {
  *nargs = -1;
  for (const R_CallMethodDef *m = shim_registered; m && m->name; ++m)
    if (strcmp(m->name, name) == 0) { *nargs = m->numArgs; return m->fun; }
  return NULL;
}

typedef SEXP (*fn9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn3)(SEXP, SEXP, SEXP);

/* .Call trampolines: return NULL when the callee raised error() */
SEXP shim_call9(fn9 f, SEXP a, SEXP b, SEXP c, SEXP d, SEXP e, SEXP g, SEXP h, SEXP i, SEXP j) {
  SEXP out = NULL;
  shim_msg[0] = 0;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) out = f(a, b, c, d, e, g, h, i, j);
  shim_jmp_armed = 0;
  return out;
}

/* generic .Call trampoline, up to 10 arguments */
typedef SEXP (*fn1)(SEXP);
typedef SEXP (*fn2)(SEXP, SEXP);
typedef SEXP (*fn8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP shim_calln(void *f, int n, SEXP *a) {
  SEXP out = NULL;
  shim_msg[0] = 0;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) {
    switch (n) {
      case 1: out = ((fn1)f)(a[0]); break;
      case 2: out = ((fn2)f)(a[0], a[1]); break;
      case 3: out = ((fn3)f)(a[0], a[1], a[2]); break;
      case 8: out = ((fn8)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
      case 9: out = ((fn9)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
      case 10: out = ((fn10)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
      case 12: out = ((fn12)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
      default: snprintf(shim_msg, sizeof shim_msg, "shim_calln: arity %d not supported", n); break;
    }
  }
  shim_jmp_armed = 0;
  return out;
}

SEXP shim_call3(fn3 f, SEXP a, SEXP b, SEXP c) {
  SEXP out = NULL;
  shim_msg[0] = 0;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) out = f(a, b, c);
  shim_jmp_armed = 0;
  return out;
}
