/* Mock of the C ABI (include/gwsem_b200.h) for `make -C src check`: records what the .Call layer hands over, so the
 * harness can compare it with what R/flatten.R (here: gwsem_b200.gwas.call_arguments) put in.  TEST ONLY. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gwsem_b200.h"
#include "mock_gwsem.h"

struct mock_record mock;
static char last_error[256] = "";

int gwsem_abi_version(void) { return GWSEM_ABI_VERSION; }
const char* gwsem_last_error(void) { return last_error; }
int gwsem_device_count(void) { return 1; }
int gwsem_engine_create(int device, gwsem_engine** out) { mock.device = device; *out = (gwsem_engine*) &mock; mock.live++; return 0; }
void gwsem_engine_destroy(gwsem_engine* e) { (void) e; mock.live--; }
int gwsem_source_open(const char* path, const char* method, int src_rows, gwsem_source** out) {
  snprintf(mock.path, sizeof(mock.path), "%s", path);
  snprintf(mock.method, sizeof(mock.method), "%s", method);
  mock.src_rows = src_rows;
  if (strstr(path, "does-not-exist")) { snprintf(last_error, sizeof(last_error), "cannot open %s", path); return 1; }
  *out = (gwsem_source*) &mock;
  mock.live++;
  return 0;
}
void gwsem_source_close(gwsem_source* s) { (void) s; mock.live--; }
int gwsem_source_num_variants(const gwsem_source* s) { (void) s; return 199; }
int gwsem_model_create(gwsem_engine* e, const gwsem_model_spec* spec, gwsem_model** out) {
  (void) e;
  mock.spec = *spec;     /* pointers stay valid for the duration of the .Call */
  *out = (gwsem_model*) &mock;
  mock.live++;
  return 0;
}
void gwsem_model_destroy(gwsem_model* m) { (void) m; mock.live--; }
int gwsem_model_set_row_filter(gwsem_model* m, const int32_t* row_filter, int src_rows) {
  (void) m;
  mock.row_filter = row_filter;
  mock.row_filter_rows = src_rows;
  return 0;
}
int gwsem_gwas_run(gwsem_model* m, gwsem_source* s, const gwsem_job* job, int64_t* rows_written) {
  (void) m; (void) s;
  mock.job = *job;
  mock.runs++;
  if (mock.check) mock.check();     /* the harness looks at the borrowed buffers while they are alive */
  if (job->n_snp > 0 && job->snp[0] > 199) { snprintf(last_error, sizeof(last_error), "out of data"); return 1; }
  *rows_written = job->n_snp >= 0 ? job->n_snp : 199 - job->start_from + 1;
  return 0;
}
