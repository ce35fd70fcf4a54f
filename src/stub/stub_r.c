/* Implementation of the stand-in SEXP API (src/stub/Rinternals.h).  TEST INFRASTRUCTURE ONLY. */
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R_ext/Rdynload.h"
#include "Rinternals.h"

static struct stub_sexp nil_value = {NILSXP, 0, NULL, NULL}, names_symbol = {NILSXP, 0, NULL, NULL};
SEXP R_NilValue = &nil_value, R_NamesSymbol = &names_symbol;
int R_NaInt = INT_MIN;
jmp_buf stub_error_jmp;
int stub_error_armed = 0;
char stub_error_message[2048];

int TYPEOF(SEXP x) { return x->type; }
int LENGTH(SEXP x) { return x->length; }
int* INTEGER(SEXP x) { return (int*) x->data; }
int* LOGICAL(SEXP x) { return (int*) x->data; }
double* REAL(SEXP x) { return (double*) x->data; }
Rbyte* RAW(SEXP x) { return (Rbyte*) x->data; }
SEXP VECTOR_ELT(SEXP x, int i) { return ((SEXP*) x->data)[i]; }
SEXP STRING_ELT(SEXP x, int i) { return ((SEXP*) x->data)[i]; }
const char* CHAR(SEXP x) { return (const char*) x->data; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
int Rf_asInteger(SEXP x) {
  if (x->length < 1) return R_NaInt;
  if (x->type == INTSXP || x->type == LGLSXP) return INTEGER(x)[0];
  if (x->type == REALSXP) return (int) REAL(x)[0];
  return R_NaInt;
}
double Rf_asReal(SEXP x) {
  if (x->length < 1) return 0.0 / 0.0;
  if (x->type == REALSXP) return REAL(x)[0];
  if (x->type == INTSXP || x->type == LGLSXP) return (double) INTEGER(x)[0];
  return 0.0 / 0.0;
}
SEXP Rf_getAttrib(SEXP x, SEXP what) { return (what == R_NamesSymbol && x->names) ? x->names : R_NilValue; }

SEXP stub_vector(int type, int n) {
  SEXP v = (SEXP) calloc(1, sizeof(*v));
  v->type = type;
  v->length = n;
  size_t el = type == REALSXP ? sizeof(double) : (type == RAWSXP ? 1 : ((type == VECSXP || type == STRSXP) ? sizeof(SEXP) : sizeof(int)));
  v->data = calloc(n > 0 ? n : 1, el);
  if (type == VECSXP || type == STRSXP)
    for (int i = 0; i < n; ++i) ((SEXP*) v->data)[i] = R_NilValue;
  return v;
}
SEXP stub_string(const char* s) {
  SEXP c = (SEXP) calloc(1, sizeof(*c));
  c->type = CHARSXP;
  c->length = (int) strlen(s);
  c->data = strdup(s);
  return c;
}
void stub_set_names(SEXP list, const char* const* names) {
  SEXP nm = stub_vector(STRSXP, list->length);
  for (int i = 0; i < list->length; ++i) ((SEXP*) nm->data)[i] = stub_string(names[i]);
  list->names = nm;
}
SEXP Rf_ScalarReal(double v) { SEXP x = stub_vector(REALSXP, 1); REAL(x)[0] = v; return x; }
SEXP Rf_ScalarInteger(int v) { SEXP x = stub_vector(INTSXP, 1); INTEGER(x)[0] = v; return x; }
char* R_alloc(size_t n, int size) { return (char*) calloc(n > 0 ? n : 1, (size_t) size); }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(stub_error_message, sizeof(stub_error_message), fmt, ap);
  va_end(ap);
  if (stub_error_armed) longjmp(stub_error_jmp, 1);
  fprintf(stderr, "Error: %s\n", stub_error_message);
  exit(2);
}
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void) c; (void) f; (void) e;
  info->call = call;
  return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { info->dynamic = value; return value; }
