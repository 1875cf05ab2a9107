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

SEXP shim_call3(fn3 f, SEXP a, SEXP b, SEXP c) {
  SEXP out = NULL;
  shim_msg[0] = 0;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) out = f(a, b, c);
  shim_jmp_armed = 0;
  return out;
}
