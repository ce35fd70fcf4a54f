/* Harness of the .Call layer (src/gwsem_call.c) for machines without R.
 *
 *   check_shim <call-dump> mock     link against check/mock_gwsem.c: every field the mock receives is compared with the
 *                                   dump, error paths (missing element, wrong length, engine error) are provoked, and
 *                                   the handle count must return to zero on every path
 *   check_shim <call-dump> real     link against libgwsem_b200.so: runs the GWAS described by the dump on the GPU and
 *                                   prints the number of rows written (tests/test_r_boundary.py compares the log with
 *                                   the one GWAS() of the Python mirror writes for the same arguments)
 *
 * The dump is written by gwsem_b200.gwas.dump_call(): the six arguments of .Call("gwsem_run", ...) exactly as
 * R/flatten.R assembles them, one value per line `<list> <name> <type> <n> v...`. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "gwsem_b200.h"
#ifdef USE_MOCK
#include "mock_gwsem.h"
#endif

SEXP gwsem_run(SEXP spec, SEXP data, SEXP path, SEXP method, SEXP rowFilter, SEXP job);
SEXP gwsem_num_records(SEXP path, SEXP method, SEXP srcRows);
void R_init_gwsem(DllInfo* dll);

#define MAXEL 64
struct builder { SEXP items[MAXEL]; char* names[MAXEL]; int n; };
static void add(struct builder* b, const char* name, SEXP v) {
  if (b->n >= MAXEL) { fprintf(stderr, "too many elements\n"); exit(2); }
  b->items[b->n] = v;
  b->names[b->n++] = strdup(name);
}
static SEXP finish(struct builder* b, int named) {
  SEXP l = stub_vector(VECSXP, b->n);
  for (int i = 0; i < b->n; ++i) ((SEXP*) l->data)[i] = b->items[i];
  if (named) stub_set_names(l, (const char* const*) b->names);
  return l;
}
static char* unquote(const char* s) {
  char* o = (char*) malloc(strlen(s) + 1);
  int k = 0;
  for (const char* c = s; *c; ++c)
    if (*c == '%' && c[1] && c[2]) { char h[3] = {c[1], c[2], 0}; o[k++] = (char) strtol(h, NULL, 16); c += 2; }
    else o[k++] = *c;
  o[k] = 0;
  return o;
}

static int fails = 0;
#define EXPECT(cond) do { if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

static SEXP g_spec, g_data, g_exo, g_job, g_path, g_method, g_rf;

#ifdef USE_MOCK
static SEXP byname(SEXP list, const char* name) {
  SEXP nm = Rf_getAttrib(list, R_NamesSymbol);
  for (int i = 0; i < LENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(nm, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}
static int same_reals(const double* a, const double* b, int n) {
  for (int i = 0; i < n; ++i)
    if (!(a[i] == b[i] || (isnan(a[i]) && isnan(b[i])))) return 0;
  return 1;
}
/* runs inside gwsem_gwas_run of the mock: every borrowed buffer is still alive */
static void compare_with_dump(void) {
  const gwsem_model_spec* s = &mock.spec;
  EXPECT(s->n_vars == Rf_asInteger(byname(g_spec, "n_vars")));
  EXPECT(s->n_manifest == Rf_asInteger(byname(g_spec, "n_manifest")));
  EXPECT(s->n_params == Rf_asInteger(byname(g_spec, "n_params")));
  EXPECT(s->n_cells == Rf_asInteger(byname(g_spec, "n_cells")));
  EXPECT(same_reals(s->cells, REAL(byname(g_spec, "cells")), 5 * s->n_cells));
  EXPECT(same_reals(s->par_start, REAL(byname(g_spec, "par_start")), s->n_params));
  EXPECT(same_reals(s->par_lbound, REAL(byname(g_spec, "par_lbound")), s->n_params));
  EXPECT(s->fit == Rf_asInteger(byname(g_spec, "fit")));
  EXPECT(s->min_variance == Rf_asReal(byname(g_spec, "min_variance")));
  EXPECT(s->n_rows == Rf_asInteger(byname(g_spec, "n_rows")));
  EXPECT(memcmp(s->manifest_kind, INTEGER(byname(g_spec, "manifest_kind")), sizeof(int) * s->n_manifest) == 0);
  EXPECT(memcmp(s->n_categories, INTEGER(byname(g_spec, "n_categories")), sizeof(int) * s->n_manifest) == 0);
  for (int j = 0; j < s->n_manifest; ++j) {
    SEXP c = VECTOR_ELT(g_data, j);
    if (Rf_isNull(c)) EXPECT(s->manifest_data[j] == NULL && s->manifest_kind[j] == 1);
    else EXPECT(s->manifest_data[j] == REAL(c));           /* borrowed, not copied */
  }
  EXPECT(s->n_exo == LENGTH(g_exo));
  for (int k = 0; k < s->n_exo; ++k) {
    SEXP c = VECTOR_ELT(g_exo, k);
    if (Rf_isNull(c)) EXPECT(s->exo_data[k] == NULL && s->exo_kind[k] == 1);
    else EXPECT(s->exo_data[k] == REAL(c));
  }
  EXPECT(memcmp(s->exo_latent, INTEGER(byname(g_spec, "exo_latent")), sizeof(int) * s->n_exo) == 0);
  EXPECT(memcmp(s->exo_kind, INTEGER(byname(g_spec, "exo_kind")), sizeof(int) * s->n_exo) == 0);
  EXPECT(memcmp(s->exo_free, RAW(byname(g_spec, "exo_free")), (size_t) s->n_manifest * s->n_exo) == 0);
  EXPECT(s->n_thresholds == Rf_asInteger(byname(g_spec, "n_thresholds")));
  EXPECT(same_reals(s->thresholds, REAL(byname(g_spec, "thresholds")), 4 * s->n_thresholds));
  const gwsem_job* j = &mock.job;
  SEXP snp = byname(g_job, "snp");
  if (Rf_isNull(snp)) EXPECT(j->n_snp == -1 && j->snp == NULL);
  else EXPECT(j->n_snp == LENGTH(snp) && j->snp == INTEGER(snp));
  EXPECT(j->start_from == Rf_asInteger(byname(g_job, "start_from")));
  EXPECT(strcmp(j->out_path, CHAR(STRING_ELT(byname(g_job, "out"), 0))) == 0);
  for (int k = 0; k < s->n_params; ++k) EXPECT(strcmp(j->par_names[k], CHAR(STRING_ELT(byname(g_job, "par_names"), k))) == 0);
  for (int k = 0; k < s->n_manifest; ++k) EXPECT(strcmp(j->manifest_names[k], CHAR(STRING_ELT(byname(g_job, "manifest_names"), k))) == 0);
  SEXP ctx = byname(g_job, "context_path");
  if (Rf_isNull(ctx)) EXPECT(j->context_path == NULL);
  else EXPECT(strcmp(j->context_path, CHAR(STRING_ELT(ctx, 0))) == 0);
  SEXP vp = byname(g_job, "vcov_pairs");
  if (Rf_isNull(vp)) EXPECT(j->n_vcov_pairs == 0);
  else EXPECT(j->n_vcov_pairs == LENGTH(vp) / 2 && j->vcov_pairs == INTEGER(vp));
  EXPECT(strcmp(j->model_name, CHAR(STRING_ELT(byname(g_job, "model_name"), 0))) == 0);
  EXPECT(mock.device == Rf_asInteger(byname(g_job, "device")));
  EXPECT(strcmp(mock.path, CHAR(STRING_ELT(g_path, 0))) == 0 && strcmp(mock.method, CHAR(STRING_ELT(g_method, 0))) == 0);
  if (Rf_isNull(g_rf)) EXPECT(mock.row_filter == NULL && mock.src_rows == s->n_rows);
  else EXPECT(mock.row_filter == LOGICAL(g_rf) && mock.src_rows == LENGTH(g_rf) && mock.row_filter_rows == LENGTH(g_rf));
}
/* calls gwsem_run expecting an R error whose message contains `needle` */
static void expect_error(SEXP spec, SEXP data, SEXP path, SEXP rf, SEXP job, const char* needle) {
  stub_error_armed = 1;
  stub_error_message[0] = 0;
  if (setjmp(stub_error_jmp) == 0) {
    gwsem_run(spec, data, path, g_method, rf, job);
    fprintf(stderr, "FAILED: expected an error containing '%s'\n", needle);
    ++fails;
  } else if (!strstr(stub_error_message, needle)) {
    fprintf(stderr, "FAILED: error '%s' does not contain '%s'\n", stub_error_message, needle);
    ++fails;
  }
  stub_error_armed = 0;
  EXPECT(mock.live == 0);      /* engine, source and model released on the error path too */
}
static SEXP without(SEXP list, const char* drop) {
  struct builder b = {{0}, {0}, 0};
  SEXP nm = Rf_getAttrib(list, R_NamesSymbol);
  for (int i = 0; i < LENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(nm, i)), drop) != 0) add(&b, CHAR(STRING_ELT(nm, i)), VECTOR_ELT(list, i));
  return finish(&b, 1);
}
static SEXP with(SEXP list, const char* name, SEXP v) {
  struct builder b = {{0}, {0}, 0};
  SEXP nm = Rf_getAttrib(list, R_NamesSymbol);
  for (int i = 0; i < LENGTH(list); ++i)
    add(&b, CHAR(STRING_ELT(nm, i)), strcmp(CHAR(STRING_ELT(nm, i)), name) == 0 ? v : VECTOR_ELT(list, i));
  return finish(&b, 1);
}
#endif

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: check_shim <call-dump> mock|real\n"); return 2; }
  FILE* f = fopen(argv[1], "r");
  if (!f) { perror(argv[1]); return 2; }
  struct builder spec = {{0}, {0}, 0}, data = {{0}, {0}, 0}, exo = {{0}, {0}, 0}, job = {{0}, {0}, 0};
  g_rf = R_NilValue;
  size_t cap = 1 << 26;
  char* line = (char*) malloc(cap);
  while (fgets(line, (int) cap, f)) {
    char *save = NULL, *lst = strtok_r(line, " \n", &save), *name = strtok_r(NULL, " \n", &save);
    char *type = strtok_r(NULL, " \n", &save), *cnt = strtok_r(NULL, " \n", &save);
    if (!lst || !name || !type || !cnt) continue;
    const int n = atoi(cnt);
    SEXP v = R_NilValue;
    if (!strcmp(type, "int") || !strcmp(type, "lgl")) {
      v = stub_vector(!strcmp(type, "int") ? INTSXP : LGLSXP, n);
      for (int i = 0; i < n; ++i) INTEGER(v)[i] = atoi(strtok_r(NULL, " \n", &save));
    } else if (!strcmp(type, "real")) {
      v = stub_vector(REALSXP, n);
      for (int i = 0; i < n; ++i) REAL(v)[i] = strtod(strtok_r(NULL, " \n", &save), NULL);
    } else if (!strcmp(type, "raw")) {
      v = stub_vector(RAWSXP, n);
      for (int i = 0; i < n; ++i) RAW(v)[i] = (Rbyte) atoi(strtok_r(NULL, " \n", &save));
    } else if (!strcmp(type, "str")) {
      v = stub_vector(STRSXP, n);
      for (int i = 0; i < n; ++i) ((SEXP*) v->data)[i] = stub_string(unquote(strtok_r(NULL, " \n", &save)));
    }
    if (!strcmp(lst, "spec")) { if (strcmp(name, "n_exo")) add(&spec, name, v); }
    else if (!strcmp(lst, "data")) add(&data, name, v);
    else if (!strcmp(lst, "exo")) add(&exo, name, v);
    else if (!strcmp(lst, "job")) add(&job, name, v);
    else if (!strcmp(lst, "top")) {
      if (!strcmp(name, "path")) g_path = v;
      else if (!strcmp(name, "method")) g_method = v;
      else if (!strcmp(name, "rowFilter")) g_rf = v;
    }
  }
  fclose(f);
  g_exo = finish(&exo, 0);
  add(&spec, "exo", g_exo);
  g_spec = finish(&spec, 1);
  g_data = finish(&data, 0);
  g_job = finish(&job, 1);

  DllInfo dll = {NULL, 1};
  R_init_gwsem(&dll);
  int n_entries = 0;
  while (dll.call && dll.call[n_entries].name) ++n_entries;
  EXPECT(n_entries == 3 && dll.dynamic == 0);
  EXPECT(strcmp(dll.call[0].name, "gwsem_run") == 0 && dll.call[0].numArgs == 6);

#ifdef USE_MOCK
  (void) argv;
  mock.check = compare_with_dump;
  SEXP rows = gwsem_run(g_spec, g_data, g_path, g_method, g_rf, g_job);
  EXPECT(mock.runs == 1 && mock.live == 0);
  SEXP snp = byname(g_job, "snp");
  EXPECT(REAL(rows)[0] == (Rf_isNull(snp) ? 199 - Rf_asInteger(byname(g_job, "start_from")) + 1 : LENGTH(snp)));
  mock.check = NULL;
  /* error paths */
  expect_error(without(g_spec, "par_start"), g_data, g_path, g_rf, g_job, "'par_start' is missing");
  expect_error(with(g_spec, "n_params", Rf_ScalarInteger(Rf_asInteger(byname(g_spec, "n_params")) + 1)), g_data, g_path, g_rf, g_job, "expected");
  expect_error(with(g_spec, "n_rows", Rf_ScalarInteger(7)), g_data, g_path, g_rf, g_job, "rows, expected 7");
  { SEXP p2 = stub_vector(STRSXP, 1); ((SEXP*) p2->data)[0] = stub_string("does-not-exist.pgen");
    expect_error(g_spec, g_data, p2, g_rf, g_job, "cannot open"); }
  { SEXP far = stub_vector(INTSXP, 1); INTEGER(far)[0] = 250;          /* tests/testthat/test-model.R:60-66 */
    expect_error(g_spec, g_data, g_path, g_rf, with(g_job, "snp", far), "out of data"); }
  { SEXP r = gwsem_num_records(g_path, g_method, Rf_ScalarInteger(-1)); EXPECT(INTEGER(r)[0] == 199 && mock.live == 0); }
  if (fails == 0) printf("check_shim: ok (marshalling of %d manifest variables, %d parameters, %d exogenous covariates; 5 error paths)\n",
                         mock.spec.n_manifest, mock.spec.n_params, mock.spec.n_exo);
#else
  SEXP rows = gwsem_run(g_spec, g_data, g_path, g_method, g_rf, g_job);
  printf("rows written: %.0f\n", REAL(rows)[0]);
#endif
  return fails ? 1 : 0;
}
