/* Minimal stand-in for R's Rinternals.h: just enough of the SEXP API for src/gwsem_call.c to be compiled and
 * exercised by `make -C src check` on a machine without R (this image has none).  TEST INFRASTRUCTURE ONLY --
 * the package itself is built against R's own headers (src/Makevars). */
#ifndef GWSEM_STUB_RINTERNALS_H_
#define GWSEM_STUB_RINTERNALS_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, RAWSXP = 24, CHARSXP = 9 };
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
typedef struct stub_sexp {
  int type, length;
  void* data;               /* int* / double* / Rbyte* / SEXP* (VECSXP, STRSXP) / char* (CHARSXP) */
  struct stub_sexp* names;  /* STRSXP or NULL */
} *SEXP;

extern SEXP R_NilValue, R_NamesSymbol;
extern int R_NaInt;
#define NA_INTEGER R_NaInt

int TYPEOF(SEXP x);
int LENGTH(SEXP x);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
double* REAL(SEXP x);
Rbyte* RAW(SEXP x);
SEXP VECTOR_ELT(SEXP x, int i);
SEXP STRING_ELT(SEXP x, int i);
const char* CHAR(SEXP x);
int Rf_isNull(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP what);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarInteger(int v);
char* R_alloc(size_t n, int size);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));

/* constructors used by the harness (not part of R's API) */
SEXP stub_vector(int type, int n);
SEXP stub_string(const char* s);
void stub_set_names(SEXP list, const char* const* names);
/* Rf_error longjmps here when armed; the message is kept */
#include <setjmp.h>
extern jmp_buf stub_error_jmp;
extern int stub_error_armed;
extern char stub_error_message[2048];

#ifdef __cplusplus
}
#endif
#endif
