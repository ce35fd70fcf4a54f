/* gwsem_call.c -- the .Call layer between the unchanged gwsem R API and libgwsem_b200.
 *
 * Replaces what the reference keeps in src/: init.cpp:7-21 (R_init_gwsem; its R_CallMethodDef
 * table is empty because the reference reaches native code through OpenMx's provider plugin,
 * src/loader.cpp:421-433) and the two providers themselves.  Here the whole per-SNP loop is
 * native, so R hands over the flattened model (R/flatten.R: flattenModel, manifestColumns,
 * jobList) in one .Call and gets the number of rows written to the checkpoint log back.
 *
 *   .Call("gwsem_run", spec, data, path, method, rowFilter, job)      GWAS(), R/model.R:305-315
 *   .Call("gwsem_num_records", path, method, srcRows)                 numAvailableRecords(), R/model.R:210-224
 *   .Call("gwsem_devices")                                           number of usable GPUs
 *
 * Only marshalling happens here: every list element is looked up by name, checked for type and
 * length, and copied or borrowed into gwsem_model_spec / gwsem_job (include/gwsem_b200.h).
 * Errors of the engine surface as R errors (Rf_error) with the engine's message, e.g.
 * "out of data" (tests/testthat/test-model.R:60-66).
 * Compiles against R's headers (R CMD INSTALL, see Makevars) and, for `make check`, against the
 * minimal stand-ins in src/stub/.
 */
#include <stdint.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Visibility.h>

#include "gwsem_b200.h"

/* ---- named-list access ------------------------------------------------------------------- */
static SEXP el(SEXP list, const char* name, int required) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (!Rf_isNull(names))
    for (int i = 0; i < LENGTH(list); ++i)
      if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  if (required) Rf_error("gwsem: list element '%s' is missing", name);
  return R_NilValue;
}
static const double* reals(SEXP list, const char* name, int n) {
  SEXP x = el(list, name, 1);
  if (TYPEOF(x) != REALSXP) Rf_error("gwsem: '%s' must be a double vector", name);
  if (n >= 0 && LENGTH(x) != n) Rf_error("gwsem: '%s' has length %d, expected %d", name, LENGTH(x), n);
  return REAL(x);
}
static const int* ints(SEXP list, const char* name, int n) {
  SEXP x = el(list, name, 1);
  if (TYPEOF(x) != INTSXP && TYPEOF(x) != LGLSXP) Rf_error("gwsem: '%s' must be an integer vector", name);
  if (n >= 0 && LENGTH(x) != n) Rf_error("gwsem: '%s' has length %d, expected %d", name, LENGTH(x), n);
  return INTEGER(x);
}
static int int1(SEXP list, const char* name) { return Rf_asInteger(el(list, name, 1)); }
static double real1(SEXP list, const char* name) { return Rf_asReal(el(list, name, 1)); }
static const char* str1(SEXP list, const char* name, int required) {
  SEXP x = el(list, name, required);
  if (Rf_isNull(x)) return NULL;
  if (TYPEOF(x) != STRSXP || LENGTH(x) != 1) Rf_error("gwsem: '%s' must be one string", name);
  return CHAR(STRING_ELT(x, 0));
}
/* columns of a list of double vectors; NULL entries stay NULL (the genotype itself) */
static const double** columns(SEXP list, int n, int n_rows, const char* what) {
  if (LENGTH(list) != n) Rf_error("gwsem: %s has %d columns, expected %d", what, LENGTH(list), n);
  const double** out = (const double**) R_alloc(n > 0 ? n : 1, sizeof(double*));
  for (int j = 0; j < n; ++j) {
    SEXP c = VECTOR_ELT(list, j);
    if (Rf_isNull(c)) { out[j] = NULL; continue; }
    if (TYPEOF(c) != REALSXP) Rf_error("gwsem: %s column %d must be a double vector", what, j + 1);
    if (LENGTH(c) != n_rows) Rf_error("gwsem: %s column %d has %d rows, expected %d", what, j + 1, LENGTH(c), n_rows);
    out[j] = REAL(c);
  }
  return out;
}
static const char** strings(SEXP x, int n, const char* what) {
  if (TYPEOF(x) != STRSXP || LENGTH(x) != n) Rf_error("gwsem: %s must be a character vector of length %d", what, n);
  const char** out = (const char**) R_alloc(n > 0 ? n : 1, sizeof(char*));
  for (int i = 0; i < n; ++i) out[i] = CHAR(STRING_ELT(x, i));
  return out;
}

/* Handles are released on every path: Rf_error does not return. */
struct handles { gwsem_engine* e; gwsem_source* s; gwsem_model* m; };
static void release(struct handles* h) {
  if (h->m) gwsem_model_destroy(h->m);
  if (h->s) gwsem_source_close(h->s);
  if (h->e) gwsem_engine_destroy(h->e);
  h->m = NULL; h->s = NULL; h->e = NULL;
}
static void chk(int rc, struct handles* h) {
  if (rc == 0) return;
  char msg[2048];
  strncpy(msg, gwsem_last_error(), sizeof(msg) - 1);
  msg[sizeof(msg) - 1] = 0;
  release(h);
  Rf_error("%s", msg);
}

/* .Call("gwsem_run", spec, data, path, method, rowFilter, job) -> rows written (double) */
SEXP gwsem_run(SEXP spec, SEXP data, SEXP path, SEXP method, SEXP rowFilter, SEXP job) {
  struct handles h = {NULL, NULL, NULL};
  if (TYPEOF(path) != STRSXP || LENGTH(path) != 1) Rf_error("gwsem: snpData must be one file name");
  if (TYPEOF(method) != STRSXP || LENGTH(method) != 1) Rf_error("gwsem: method must be one string");

  gwsem_model_spec ms;
  memset(&ms, 0, sizeof(ms));
  ms.n_vars = int1(spec, "n_vars");
  ms.n_manifest = int1(spec, "n_manifest");
  ms.n_params = int1(spec, "n_params");
  ms.n_cells = int1(spec, "n_cells");
  ms.cells = reals(spec, "cells", 5 * ms.n_cells);            /* row major: matrix, row, col, parameter, value */
  ms.par_start = reals(spec, "par_start", ms.n_params);
  ms.par_lbound = reals(spec, "par_lbound", ms.n_params);     /* NA = unbounded (NA_real_ is a NaN) */
  ms.fit = int1(spec, "fit");
  ms.min_variance = real1(spec, "min_variance");
  ms.n_rows = int1(spec, "n_rows");
  ms.manifest_kind = ints(spec, "manifest_kind", ms.n_manifest);
  ms.manifest_data = columns(data, ms.n_manifest, ms.n_rows, "data");
  ms.n_categories = ints(spec, "n_categories", ms.n_manifest);
  SEXP exo = el(spec, "exo", 1);
  ms.n_exo = LENGTH(exo);
  ms.exo_data = columns(exo, ms.n_exo, ms.n_rows, "exo");
  ms.exo_latent = ints(spec, "exo_latent", ms.n_exo);
  ms.exo_kind = ints(spec, "exo_kind", ms.n_exo);
  SEXP xf = el(spec, "exo_free", 1);                          /* n_manifest x n_exo, row major */
  if (TYPEOF(xf) != RAWSXP || LENGTH(xf) != ms.n_manifest * ms.n_exo) Rf_error("gwsem: 'exo_free' must be raw(n_manifest * n_exo)");
  ms.exo_free = RAW(xf);
  ms.n_thresholds = int1(spec, "n_thresholds");
  ms.thresholds = reals(spec, "thresholds", 4 * ms.n_thresholds);   /* row major: manifest, k, parameter, value */

  int src_rows = ms.n_rows;
  const int* rf = NULL;
  if (!Rf_isNull(rowFilter)) {                                /* TRUE = skip the row (src/LoadDataAPI.h:80) */
    if (TYPEOF(rowFilter) != LGLSXP) Rf_error("gwsem: rowFilter must be a logical vector");
    rf = LOGICAL(rowFilter);
    src_rows = LENGTH(rowFilter);
  }

  gwsem_job jb;
  memset(&jb, 0, sizeof(jb));
  SEXP snp = el(job, "snp", 0);
  if (Rf_isNull(snp)) { jb.snp = NULL; jb.n_snp = -1; }       /* every SNP from start_from */
  else {
    if (TYPEOF(snp) != INTSXP) Rf_error("gwsem: SNP must be an integer vector");
    jb.snp = INTEGER(snp);
    jb.n_snp = LENGTH(snp);
  }
  jb.start_from = int1(job, "start_from");
  jb.out_path = str1(job, "out", 1);
  jb.par_names = strings(el(job, "par_names", 1), ms.n_params, "par_names");
  jb.context_path = str1(job, "context_path", 0);             /* .pvar / .bim; NULL for bgen */
  SEXP vp = el(job, "vcov_pairs", 0);
  if (!Rf_isNull(vp)) {
    if (TYPEOF(vp) != INTSXP || LENGTH(vp) % 2) Rf_error("gwsem: vcov_pairs must hold pairs of 0-based parameter indices");
    jb.vcov_pairs = INTEGER(vp);
    jb.n_vcov_pairs = LENGTH(vp) / 2;
  }
  jb.batch_snps = Rf_isNull(el(job, "batch_snps", 0)) ? 0 : int1(job, "batch_snps");
  jb.model_name = str1(job, "model_name", 1);
  jb.manifest_names = strings(el(job, "manifest_names", 1), ms.n_manifest, "manifest_names");
  const int device = Rf_isNull(el(job, "device", 0)) ? 0 : int1(job, "device");

  chk(gwsem_engine_create(device, &h.e), &h);
  chk(gwsem_source_open(CHAR(STRING_ELT(path, 0)), CHAR(STRING_ELT(method, 0)), src_rows, &h.s), &h);
  chk(gwsem_model_create(h.e, &ms, &h.m), &h);
  chk(gwsem_model_set_row_filter(h.m, rf, src_rows), &h);
  int64_t rows = 0;
  chk(gwsem_gwas_run(h.m, h.s, &jb, &rows), &h);
  release(&h);
  return Rf_ScalarReal((double) rows);
}

/* .Call("gwsem_num_records", path, method, srcRows) -> integer (NA when the file does not say, i.e. a .bed
 * without its sample count: R/model.R:216-223, tests/testthat/test-formats.R:25-33) */
SEXP gwsem_num_records(SEXP path, SEXP method, SEXP srcRows) {
  struct handles h = {NULL, NULL, NULL};
  if (TYPEOF(path) != STRSXP || LENGTH(path) != 1) Rf_error("gwsem: path must be one file name");
  chk(gwsem_source_open(CHAR(STRING_ELT(path, 0)), CHAR(STRING_ELT(method, 0)), Rf_asInteger(srcRows), &h.s), &h);
  const int n = gwsem_source_num_variants(h.s);
  release(&h);
  return Rf_ScalarInteger(n < 0 ? NA_INTEGER : n);
}

SEXP gwsem_devices(void) { return Rf_ScalarInteger(gwsem_device_count()); }

static const R_CallMethodDef CallEntries[] = {
  {"gwsem_run", (DL_FUNC) &gwsem_run, 6},
  {"gwsem_num_records", (DL_FUNC) &gwsem_num_records, 3},
  {"gwsem_devices", (DL_FUNC) &gwsem_devices, 0},
  {NULL, NULL, 0}
};

void attribute_visible R_init_gwsem(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
  if (gwsem_abi_version() != GWSEM_ABI_VERSION)
    Rf_error("gwsem: libgwsem_b200 has ABI version %d, this package was built for %d", gwsem_abi_version(), GWSEM_ABI_VERSION);
}
