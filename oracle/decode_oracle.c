/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's genotype decode path.
 *
 * Plain-C, scalar, one-sample-at-a-time restatement of what gwsem's two data
 * providers hand to OpenMx for one SNP.  It exists so that the CUDA decode
 * kernels can be checked bit-for-bit; it is never linked into, imported by, or
 * called from the product (gwsem_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_decode.py checks every function here
 * against oracle/_ref (the reference's own pgenlib compiled from
 * /root/reference/src) on all bundled fixtures and on synthetic files covering
 * every record type, and against the reference tests' numeric pins
 * (tests/testthat/test-formats.R:110-117).
 *
 * What each part follows (paths relative to /root/reference):
 *   .pgen/.bed container      src/pgenlib_misc.h:614-698 (layout), src/pgenlib_read.cc:681-977
 *                             (PgfiInitPhase1), 999-1677 (PgfiInitPhase2: vrtype / record length arrays)
 *   main genotype track       src/pgenlib_read.cc:2519-2535 (2-bit), 2428-2516 (1-bit + difflist),
 *                             2789-2840 (difflist-on-constant, all-ref), 2793-2809 + 1679-1726 (LD)
 *   difflist / deltalist      src/pgenlib_misc.h:699-718, src/pgenlib_read.cc:2008-2042, 2325-2372
 *   varint                    src/plink2_base.h:3079-3100
 *   plink1 .bed recode        src/pgenlib_read.cc:1979-1992
 *   hardcall-phase skip       src/pgenlib_read.cc:6706-6725 (SkipAux2)
 *   dosage track              src/pgenlib_read.cc:7065-7318 (ParseDosage16), vrtype bits 5-6
 *   doubles + rowFilter + MAF src/loader.cpp:211-273, 386-394
 *   bgen container            src/bgen.cpp:73-115 (header), 341-386 (block),
 *                             src/include/genfile/bgen/bgen.hpp:479-541 (variant ids)
 *   bgen layout 1 probs       bgen.hpp:783-811, src/bgen.cpp:302-315 (u16/32768)
 *   bgen layout 2 probs       bgen.hpp:1025-1052 (block header), 815-885 (bit parser), 1139-1224
 *   bgen -> dosage            src/loader.cpp:38-50 (2*p0 + p1, exact 0 -> NA), 141-143 (MAF)
 *   bgen SNP order            src/bgen.cpp:561-566, 649-663 (.bgi ORDER BY chromosome, position,
 *                             rsid, allele1, allele2)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static __thread char g_err[512];
const char* orc_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof(g_err), __VA_ARGS__); return -1; } while (0)

/* R's NA_real_ (NaN, low word 1954) */
static double na_real(void) {
  const uint64_t bits = 0x7FF00000000007A2ULL;
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

static int slurp(const char* path, uint8_t** buf, uint64_t* len) {
  FILE* f = fopen(path, "rb");
  if (!f) FAIL("cannot open %s", path);
  fseeko(f, 0, SEEK_END);
  *len = (uint64_t)ftello(f);
  rewind(f);
  *buf = (uint8_t*)malloc(*len + 16);
  if (!*buf) { fclose(f); FAIL("out of memory"); }
  if (*len && fread(*buf, 1, *len, f) != *len) { fclose(f); free(*buf); FAIL("short read on %s", path); }
  memset(*buf + *len, 0, 16);
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------ pgen / bed */

typedef struct {
  uint8_t* data;
  uint64_t len;
  uint32_t M, N;
  int mode;          /* 0x01 bed, 0x02/0x03/0x04 fixed width, 0x10 variable */
  uint8_t* vrtype;   /* per variant */
  uint64_t* fpos;    /* M+1 record offsets */
} orc_pgen;

static uint32_t vint(const uint8_t** p, const uint8_t* end) {
  uint32_t v = 0, shift = 0;
  while (*p < end) {
    uint32_t b = *(*p)++;
    v |= (b & 127u) << shift;
    if (b <= 127) return v;
    shift += 7;
  }
  return 0x80000000u;
}

static uint32_t id_bytes(uint32_t n) { /* bytes needed to hold the value n (n >= 1) */
  uint32_t b = 1;
  while (b < 4 && (n >> (8 * b))) ++b;
  return b;
}

void orc_pgen_close(orc_pgen* h) {
  if (!h) return;
  free(h->data); free(h->vrtype); free(h->fpos); free(h);
}

int orc_pgen_open(const char* path, int sample_ct_hint, orc_pgen** out) {
  orc_pgen* h = (orc_pgen*)calloc(1, sizeof(orc_pgen));
  if (slurp(path, &h->data, &h->len)) { free(h); return -1; }
  const uint8_t* d = h->data;
  if (h->len < 3 || d[0] != 0x6c || d[1] != 0x1b) { orc_pgen_close(h); FAIL("%s: bad magic", path); }
  h->mode = d[2];
  if (h->mode == 0x01) {
    if (sample_ct_hint <= 0) { orc_pgen_close(h); FAIL("%s: .bed needs a sample count", path); }
    h->N = (uint32_t)sample_ct_hint;
    uint64_t w = (h->N + 3) / 4;
    if ((h->len - 3) % w) { orc_pgen_close(h); FAIL("%s: unexpected .bed size", path); }
    h->M = (uint32_t)((h->len - 3) / w);
    h->vrtype = (uint8_t*)calloc(h->M + 1, 1);
    h->fpos = (uint64_t*)malloc(sizeof(uint64_t) * (h->M + 1));
    for (uint32_t i = 0; i <= h->M; ++i) h->fpos[i] = 3 + w * i;
    *out = h;
    return 0;
  }
  if (h->len < 12) { orc_pgen_close(h); FAIL("%s: truncated header", path); }
  memcpy(&h->M, d + 3, 4);
  memcpy(&h->N, d + 7, 4);
  const uint32_t ctrl = d[11];
  if (sample_ct_hint > 0 && (uint32_t)sample_ct_hint != h->N) {
    orc_pgen_close(h); FAIL("%s: has %u samples, caller expected %d", path, h->N, sample_ct_hint);
  }
  h->vrtype = (uint8_t*)calloc(h->M + 1, 1);
  h->fpos = (uint64_t*)malloc(sizeof(uint64_t) * (h->M + 1));
  uint64_t pos = 12;
  if (h->mode >= 0x02 && h->mode <= 0x04) {
    if ((ctrl >> 6) == 3) pos += (h->M + 7) / 8; /* explicit nonref flags */
    uint64_t w = (h->N + 3) / 4;
    uint8_t vt = 0;
    if (h->mode == 0x03) { w += 2ull * h->N; vt = 0x40; }
    if (h->mode == 0x04) { w += 4ull * h->N; vt = 0xc0; }
    if (pos + w * h->M != h->len) { orc_pgen_close(h); FAIL("%s: unexpected size", path); }
    for (uint32_t i = 0; i <= h->M; ++i) { h->fpos[i] = pos + w * i; h->vrtype[i] = vt; }
    *out = h;
    return 0;
  }
  if (h->mode != 0x10) { orc_pgen_close(h); FAIL("%s: storage mode 0x%x unsupported", path, h->mode); }
  if (ctrl & 8) { orc_pgen_close(h); FAIL("%s: fused vrtype/length headers unsupported", path); }
  const uint32_t alt_ct_bytes = (ctrl >> 4) & 3;
  if (alt_ct_bytes) { orc_pgen_close(h); FAIL("%s: multiallelic files unsupported", path); }
  const uint32_t nonref_stored = ((ctrl >> 6) == 3);
  const uint32_t wide_types = (ctrl & 4) != 0;
  const uint32_t len_bytes = 1 + (ctrl & 3);
  const uint32_t vblock_ct = (h->M + 65535) / 65536;
  uint64_t cur;
  memcpy(&cur, d + pos, 8);
  pos += 8ull * vblock_ct;
  uint32_t v = 0;
  for (uint32_t b = 0; b < vblock_ct; ++b) {
    const uint32_t cnt = (b + 1 == vblock_ct) ? h->M - 65536u * b : 65536u;
    if (wide_types) {
      memcpy(h->vrtype + v, d + pos, cnt);
      pos += cnt;
    } else {
      for (uint32_t k = 0; k < cnt; ++k) h->vrtype[v + k] = (d[pos + k / 2] >> (4 * (k & 1))) & 15;
      pos += (cnt + 1) / 2;
    }
    for (uint32_t k = 0; k < cnt; ++k) {
      uint32_t rl = 0;
      memcpy(&rl, d + pos + (uint64_t)k * len_bytes, len_bytes);
      h->fpos[v + k] = cur;
      cur += rl;
    }
    pos += (uint64_t)cnt * len_bytes;
    if (nonref_stored) pos += (cnt + 7) / 8;
    v += cnt;
  }
  h->fpos[h->M] = cur;
  if (cur > h->len) { orc_pgen_close(h); FAIL("%s: records run past end of file", path); }
  *out = h;
  return 0;
}

int orc_pgen_variant_ct(const orc_pgen* h) { return (int)h->M; }
int orc_pgen_sample_ct(const orc_pgen* h) { return (int)h->N; }
int orc_pgen_vrtype(const orc_pgen* h, int i) { return h->mode == 0x01 ? 0x100 : h->vrtype[i]; }
uint64_t orc_pgen_record_offset(const orc_pgen* h, int i) { return h->fpos[i]; }

/* Walk a difflist (with_geno=1) or a deltalist (with_geno=0).  ids[] receives the
 * sample ids, geno[] the 2-bit values.  Returns the entry count or -1. */
static int walk_list(const uint8_t** pp, const uint8_t* end, uint32_t N, int with_geno,
                     uint32_t* ids, uint8_t* geno) {
  const uint32_t len = vint(pp, end);
  if (!len) return 0;
  if (len > N / 8) FAIL("difflist longer than N/8");
  const uint32_t groups = (len + 63) / 64;
  const uint32_t ib = id_bytes(N);
  const uint8_t* group_ids = *pp;
  *pp += (uint64_t)groups * ib + (groups - 1);
  const uint8_t* packed = *pp;
  if (with_geno) *pp += (len + 3) / 4;
  if (*pp > end) FAIL("difflist header past record end");
  uint32_t sid = 0;
  for (uint32_t k = 0; k < len; ++k) {
    if ((k & 63) == 0) {
      sid = 0;
      memcpy(&sid, group_ids + (uint64_t)(k >> 6) * ib, ib);
    } else {
      sid += vint(pp, end);
    }
    if (sid >= N) FAIL("difflist sample id out of range");
    ids[k] = sid;
    if (with_geno) geno[k] = (packed[k >> 2] >> (2 * (k & 3))) & 3;
  }
  return (int)len;
}

static int apply_difflist(const uint8_t** pp, const uint8_t* end, uint32_t N, uint8_t* codes) {
  uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (N / 8 + 1));
  uint8_t* gv = (uint8_t*)malloc(N / 8 + 1);
  int n = walk_list(pp, end, N, 1, ids, gv);
  for (int k = 0; k < n; ++k) codes[ids[k]] = gv[k];
  free(ids); free(gv);
  return n < 0 ? -1 : 0;
}

/* Main-track decode of variant vidx into codes[N] (0,1,2 = ALT count, 3 = missing);
 * *after receives the read position following the main track. */
static int main_track(const orc_pgen* h, uint32_t vidx, uint8_t* codes, const uint8_t** after) {
  const uint32_t N = h->N;
  const uint8_t* p = h->data + h->fpos[vidx];
  const uint8_t* end = h->data + h->fpos[vidx + 1];
  if (h->mode == 0x01) { /* plink1: 00->2, 01->missing, 10->1, 11->0 */
    static const uint8_t recode[4] = {2, 3, 1, 0};
    for (uint32_t s = 0; s < N; ++s) codes[s] = recode[(p[s >> 2] >> (2 * (s & 3))) & 3];
    *after = p + (N + 3) / 4;
    return 0;
  }
  const uint32_t mt = h->vrtype[vidx] & 7;
  if (mt == 2 || mt == 3) { /* LD: copy the most recent non-LD variant, patch, maybe invert */
    uint32_t base = vidx;
    do {
      if (!base) FAIL("LD-compressed record %u has no base", vidx);
      --base;
    } while ((h->vrtype[base] & 6) == 2);
    const uint8_t* ignore;
    if (main_track(h, base, codes, &ignore)) return -1;
    if (apply_difflist(&p, end, N, codes)) return -1;
    if (mt == 3)
      for (uint32_t s = 0; s < N; ++s)
        if (codes[s] != 3) codes[s] = 2 - codes[s];
  } else if (mt == 0) {
    for (uint32_t s = 0; s < N; ++s) codes[s] = (p[s >> 2] >> (2 * (s & 3))) & 3;
    p += (N + 3) / 4;
  } else if (mt == 1) {
    const uint32_t c2 = *p++;
    const uint8_t lo = (c2 >> 2) & 3, hi = lo + (c2 & 3);
    for (uint32_t s = 0; s < N; ++s) codes[s] = ((p[s >> 3] >> (s & 7)) & 1) ? hi : lo;
    p += (N + 7) / 8;
    if (apply_difflist(&p, end, N, codes)) return -1;
  } else if (mt == 5) {
    memset(codes, 0, N);
  } else { /* 4, 6, 7: constant + difflist */
    memset(codes, mt & 3, N);
    if (apply_difflist(&p, end, N, codes)) return -1;
  }
  if (p > end) FAIL("record %u main track overruns", vidx);
  *after = p;
  return 0;
}

/* Full PgrGet1D equivalent: codes[N] plus dosage[N] (0xFFFF = no dosage entry). */
int orc_pgen_get(const orc_pgen* h, int vidx, uint8_t* codes, uint16_t* dosage) {
  const uint32_t N = h->N;
  if (vidx < 0 || (uint32_t)vidx >= h->M) FAIL("out of data (record %d requested but only %u in file)", vidx + 1, h->M);
  const uint8_t* p;
  if (main_track(h, (uint32_t)vidx, codes, &p)) return -1;
  for (uint32_t s = 0; s < N; ++s) dosage[s] = 0xFFFF;
  if (h->mode == 0x01) return 0;
  const uint32_t vt = h->vrtype[vidx];
  const uint8_t* end = h->data + h->fpos[vidx + 1];
  if (vt & 8) FAIL("multiallelic record %d unsupported", vidx);
  if (!(vt & 0x60)) return 0;
  if (vt & 0x10) { /* skip hardcall-phase track */
    uint32_t het = 0;
    for (uint32_t s = 0; s < N; ++s) het += (codes[s] == 1);
    const uint32_t first = 1 + het / 8;
    if (!het) FAIL("read_dosages(varient %d) error code 6", vidx); /* pgenlib_read.cc:6631-6635 */
    if (p[0] & 1) {
      uint32_t pc = 0;
      for (uint32_t k = 0; k < first; ++k) pc += (uint32_t)__builtin_popcount(p[k]);
      if (pc < 2) FAIL("read_dosages(varient %d) error code 6", vidx); /* pgenlib_read.cc:6673-6677 */
      p += first + (pc - 1 + 7) / 8;
    } else {
      p += first;
    }
  }
  const uint32_t dt = vt & 0x60;
  if (dt == 0x40) {
    for (uint32_t s = 0; s < N; ++s) {
      uint16_t v;
      memcpy(&v, p + 2ull * s, 2);
      dosage[s] = v; /* 65535 == missing == "no entry" */
    }
    p += 2ull * N;
  } else if (dt == 0x60) {
    const uint8_t* bits = p;
    p += (N + 7) / 8;
    for (uint32_t s = 0; s < N; ++s)
      if ((bits[s >> 3] >> (s & 7)) & 1) { memcpy(&dosage[s], p, 2); p += 2; }
  } else {
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (N / 8 + 1));
    int n = walk_list(&p, end, N, 0, ids, NULL);
    if (n < 0) { free(ids); return -1; }
    for (int k = 0; k < n; ++k) { memcpy(&dosage[ids[k]], p, 2); p += 2; }
    free(ids);
  }
  if (p > end) FAIL("record %d dosage track overruns", vidx);
  return 0;
}

/* loader.cpp:211-273 + 386-394: doubles with rowFilter compaction, then MAF. */
static int finish_row(uint32_t N, const uint8_t* codes, const uint16_t* dosage, const int* row_filter,
                      double* out, double* maf) {
  const double lut[4] = {0.0, 1.0, 2.0, na_real()};
  int r = 0;
  for (uint32_t s = 0; s < N; ++s) {
    if (row_filter && row_filter[s]) continue;
    out[r++] = (dosage && dosage[s] != 0xFFFF) ? (double)dosage[s] * 0.00006103515625 : lut[codes[s]];
  }
  double tot = 0;
  int cnt = 0;
  for (int k = 0; k < r; ++k)
    if (isfinite(out[k])) { tot += out[k]; ++cnt; }
  double m = tot / (2.0 * cnt);
  if (m > 0.5) m = 1 - m;
  if (maf) *maf = m;
  return r;
}

int orc_pgen_load_row(const orc_pgen* h, int vidx, const int* row_filter, double* out, double* maf) {
  uint8_t* codes = (uint8_t*)malloc(h->N);
  uint16_t* dosage = (uint16_t*)malloc(2ull * h->N);
  int rc = orc_pgen_get(h, vidx, codes, dosage);
  if (!rc) rc = finish_row(h->N, codes, dosage, row_filter, out, maf);
  free(codes); free(dosage);
  return rc;
}

/* ------------------------------------------------------------------------ bgen */

typedef struct {
  uint64_t start;   /* file offset of the variant identifying block */
  uint64_t geno;    /* file offset of the genotype data block */
  uint64_t next;
  char *snpid, *rsid, *chrom, *a1, *a2;
  uint32_t pos;
} orc_bgen_var;

typedef struct {
  uint8_t* data;
  uint64_t len;
  uint32_t M, N, flags, layout, compression;
  orc_bgen_var* var;   /* file order */
  uint32_t* order;     /* index order -> file order */
} orc_bgen;

static char* take_str(const uint8_t** p, uint32_t lenbytes) {
  uint32_t n = 0;
  memcpy(&n, *p, lenbytes);
  *p += lenbytes;
  char* s = (char*)malloc(n + 1);
  memcpy(s, *p, n);
  s[n] = 0;
  *p += n;
  return s;
}

static const orc_bgen* g_sort_ctx;
static int cmp_index_order(const void* a, const void* b) {
  const orc_bgen_var* x = &g_sort_ctx->var[*(const uint32_t*)a];
  const orc_bgen_var* y = &g_sort_ctx->var[*(const uint32_t*)b];
  int c = strcmp(x->chrom, y->chrom);
  if (c) return c;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  if ((c = strcmp(x->rsid, y->rsid))) return c;
  if ((c = strcmp(x->a1, y->a1))) return c;
  if ((c = strcmp(x->a2, y->a2))) return c;
  return x->start < y->start ? -1 : (x->start > y->start);
}

void orc_bgen_close(orc_bgen* h) {
  if (!h) return;
  if (h->var)
    for (uint32_t i = 0; i < h->M; ++i) {
      free(h->var[i].snpid); free(h->var[i].rsid); free(h->var[i].chrom);
      free(h->var[i].a1); free(h->var[i].a2);
    }
  free(h->var); free(h->order); free(h->data); free(h);
}

int orc_bgen_open(const char* path, orc_bgen** out) {
  orc_bgen* h = (orc_bgen*)calloc(1, sizeof(orc_bgen));
  if (slurp(path, &h->data, &h->len)) { free(h); return -1; }
  const uint8_t* d = h->data;
  uint32_t offset, hdr;
  memcpy(&offset, d, 4);
  memcpy(&hdr, d + 4, 4);
  memcpy(&h->M, d + 8, 4);
  memcpy(&h->N, d + 12, 4);
  if (memcmp(d + 16, "bgen", 4) && memcmp(d + 16, "\0\0\0\0", 4)) { orc_bgen_close(h); FAIL("%s: bad magic", path); }
  memcpy(&h->flags, d + 4 + hdr - 4, 4);
  h->compression = h->flags & 3;
  h->layout = (h->flags >> 2) & 15;
  if (h->layout != 1 && h->layout != 2) { orc_bgen_close(h); FAIL("%s: layout %u unsupported", path, h->layout); }
  h->var = (orc_bgen_var*)calloc(h->M, sizeof(orc_bgen_var));
  h->order = (uint32_t*)malloc(sizeof(uint32_t) * h->M);
  const uint8_t* p = d + 4 + offset;
  for (uint32_t i = 0; i < h->M; ++i) {
    orc_bgen_var* v = &h->var[i];
    v->start = (uint64_t)(p - d);
    if (h->layout == 1) p += 4; /* per-variant sample count */
    v->snpid = take_str(&p, 2);
    v->rsid = take_str(&p, 2);
    v->chrom = take_str(&p, 2);
    memcpy(&v->pos, p, 4);
    p += 4;
    uint32_t K = 2;
    if (h->layout == 2) { K = 0; memcpy(&K, p, 2); p += 2; }
    if (K != 2) { orc_bgen_close(h); FAIL("%s: variant %u is not biallelic", path, i); }
    v->a1 = take_str(&p, 4);
    v->a2 = take_str(&p, 4);
    v->geno = (uint64_t)(p - d);
    uint32_t payload;
    if (h->layout == 2 || h->compression) { memcpy(&payload, p, 4); p += 4 + payload; }
    else p += 6ull * h->N;
    v->next = (uint64_t)(p - d);
    h->order[i] = i;
    if (v->next > h->len) { orc_bgen_close(h); FAIL("%s: variant %u overruns file", path, i); }
  }
  g_sort_ctx = h;
  qsort(h->order, h->M, sizeof(uint32_t), cmp_index_order);
  *out = h;
  return 0;
}

int orc_bgen_variant_ct(const orc_bgen* h) { return (int)h->M; }
int orc_bgen_sample_ct(const orc_bgen* h) { return (int)h->N; }
int orc_bgen_layout(const orc_bgen* h) { return (int)h->layout; }
int orc_bgen_file_index(const orc_bgen* h, int index) { return (int)h->order[index]; }
uint64_t orc_bgen_file_offset(const orc_bgen* h, int index) { return h->var[h->order[index]].start; }
/* which: 0 SNPID, 1 RSID, 2 chromosome, 3 first allele, 4 second allele */
const char* orc_bgen_field(const orc_bgen* h, int index, int which) {
  const orc_bgen_var* v = &h->var[h->order[index]];
  switch (which) { case 0: return v->snpid; case 1: return v->rsid; case 2: return v->chrom;
                   case 3: return v->a1; default: return v->a2; }
}
uint32_t orc_bgen_position(const orc_bgen* h, int index) { return h->var[h->order[index]].pos; }

typedef size_t (*zstd_fn)(void*, size_t, const void*, size_t);
static int inflate_block(const orc_bgen* h, const uint8_t* src, uint32_t srclen, uint8_t* dst, uint32_t dstlen) {
  if (h->compression == 1) {
    uLongf got = dstlen;
    if (uncompress(dst, &got, src, srclen) != Z_OK || got != dstlen) FAIL("zlib inflate failed");
    return 0;
  }
  if (h->compression == 2) {
    static zstd_fn fn;
    if (!fn) {
      void* lib = dlopen("libzstd.so.1", RTLD_NOW);
      if (lib) fn = (zstd_fn)dlsym(lib, "ZSTD_decompress");
      if (!fn) FAIL("zstd runtime not available");
    }
    if (fn(dst, dstlen, src, srclen) != dstlen) FAIL("zstd inflate failed");
    return 0;
  }
  FAIL("unknown compression");
}

/* Uncompressed probability block of SNP `index` (index order). *buf is malloc'd. */
int orc_bgen_block(const orc_bgen* h, int index, uint8_t** buf, uint32_t* len) {
  if (index < 0 || (uint32_t)index >= h->M) FAIL("bgen has no more variants");
  const orc_bgen_var* v = &h->var[h->order[index]];
  const uint8_t* p = h->data + v->geno;
  if (h->layout == 1) {
    *len = 6 * h->N;
    *buf = (uint8_t*)malloc(*len + 8);
    if (h->compression) {
      uint32_t c;
      memcpy(&c, p, 4);
      return inflate_block(h, p + 4, c, *buf, *len);
    }
    memcpy(*buf, p, *len);
    return 0;
  }
  uint32_t total;
  memcpy(&total, p, 4);
  p += 4;
  if (h->compression) {
    memcpy(len, p, 4);
    *buf = (uint8_t*)calloc(*len + 8, 1);
    return inflate_block(h, p + 4, total - 4, *buf, *len);
  }
  *len = total;
  *buf = (uint8_t*)calloc(*len + 8, 1);
  memcpy(*buf, p, total);
  return 0;
}

/* One SNP column exactly as BgenXfer produces it (loader.cpp:38-50). */
int orc_bgen_load_row(const orc_bgen* h, int index, const int* row_filter, double* out, double* maf) {
  uint8_t* blk;
  uint32_t len;
  if (orc_bgen_block(h, index, &blk, &len)) return -1;
  const uint32_t N = h->N;
  int r = 0, cnt = 0;
  double tot = 0;
  int rc = 0;
  if (h->layout == 1) {
    for (uint32_t s = 0; s < N; ++s) {
      if (row_filter && row_filter[s]) continue;
      uint16_t q[3];
      memcpy(q, blk + 6ull * s, 6);
      double p0 = (double)q[0], p1 = (double)q[1];
      p0 /= 32768.0; p1 /= 32768.0;
      double dose = p1 * 1 + p0 * 2;
      if (dose == 0.0) dose = na_real(); else { tot += dose; ++cnt; }
      out[r++] = dose;
    }
  } else {
    uint32_t n2, K = 0;
    memcpy(&n2, blk, 4);
    memcpy(&K, blk + 4, 2);
    const uint32_t pmin = blk[6], pmax = blk[7];
    const uint8_t* ploidy = blk + 8;
    const uint32_t phased = blk[8 + N] & 1, bits = blk[9 + N];
    const uint8_t* probs = blk + 10 + N;
    if (n2 != N || K != 2) { snprintf(g_err, sizeof g_err, "bgen block header mismatch"); rc = -1; }
    else if (pmin != 2 || pmax != 2 || phased) {
      snprintf(g_err, sizeof g_err, "set_min_max_ploidy %u %u %u %u, not implemented", pmin, pmax,
               phased ? 4u : 3u, phased ? 4u : 3u);
      rc = -1;
    } else {
      const uint64_t mask = (~0ull) >> (64 - bits);
      const double denom = (double)mask;
      uint64_t bitpos = 0;
      for (uint32_t s = 0; s < N && !rc; ++s) {
        uint64_t v[2];
        for (int k = 0; k < 2; ++k) {
          uint64_t w = 0;
          memcpy(&w, probs + (bitpos >> 3), 8); /* block buffer is padded by 8 bytes */
          v[k] = (w >> (bitpos & 7)) & mask;
          bitpos += bits;
        }
        if (row_filter && row_filter[s]) continue;
        if (ploidy[s] & 0x80) { snprintf(g_err, sizeof g_err, "not implemented"); rc = -1; break; }
        const double p0 = (double)v[0] / denom, p1 = (double)v[1] / denom;
        double dose = p1 * 1 + p0 * 2;
        if (dose == 0.0) dose = na_real(); else { tot += dose; ++cnt; }
        out[r++] = dose;
      }
    }
  }
  free(blk);
  if (rc) return rc;
  double m = tot / (2.0 * cnt);
  if (m > .5) m = 1 - m;
  if (maf) *maf = m;
  return r;
}
