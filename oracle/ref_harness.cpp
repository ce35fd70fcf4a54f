// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C driver around the *reference's own* vendored pgenlib, compiled in
// place from /root/reference/src by oracle/Makefile into
// oracle/_ref/libgwsem_ref.so.  It performs the same call sequence that
// LoadDataPGENProvider2::loadRowImpl performs (reference src/loader.cpp:290-406)
// so the output of this library *is* the reference's decode result:
//   open   : PreinitPgfi -> PgfiInitPhase1(fname, UINT32_MAX, N, mmap=0) ->
//            PgfiInitPhase2(ctrl,1,1,0,0,M) -> PgrInit            (loader.cpp:293-363)
//   load   : PgrGet1D(allele_idx=1) -> 2-bit LUT {0,1,2,NA} with rowFilter
//            compaction -> dosage overlay value/16384 -> MAF      (loader.cpp:211-273, 374-394)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "pgenlib_read.h"

// The reference patches pgenlib's exit()/fputs() into mxThrow (plink2_base.h:116-118);
// OpenMx normally supplies it (src/openmx.cpp:41-50).
void mxThrow(const char* msg, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, msg);
  vsnprintf(buf, sizeof(buf), msg, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}

namespace {

// R's NA_real_: quiet NaN whose low word is 1954.
inline double r_na_real() {
  const uint64_t bits = 0x7FF00000000007A2ULL;
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

struct RefPgen {
  plink2::PgenFileInfo info;
  plink2::PgenReader reader;
  unsigned char* info_arena = nullptr;
  unsigned char* reader_arena = nullptr;
  uintptr_t* include_vec = nullptr;
  uintptr_t* genovec = nullptr;
  uintptr_t* dosage_present = nullptr;
  uint16_t* dosage_main = nullptr;
  plink2::PgrSampleSubsetIndex pssi;
  bool reader_live = false;
  bool info_live = false;
};

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_pgen_close(void* hv);

// sample_ct: number of people (needed for .bed, which has no header count);
// pass -1 to take it from the .pgen header.
void* ref_pgen_open(const char* path, int sample_ct) {
  using namespace plink2;
  RefPgen* h = new RefPgen;
  try {
    PreinitPgfi(&h->info);
    h->info.vrtypes = nullptr;
    PgenHeaderCtrl ctrl;
    uintptr_t info_cachelines = 0;
    char errbuf[kPglErrstrBufBlen];
    uint32_t want_samples = (sample_ct < 0) ? UINT32_MAX : uint32_t(sample_ct);
    if (PgfiInitPhase1(path, UINT32_MAX, want_samples, 0, &ctrl, &h->info,
                       &info_cachelines, errbuf) != kPglRetSuccess) {
      throw std::runtime_error(std::string("PgfiInitPhase1: ") + errbuf);
    }
    h->info_live = true;
    if (info_cachelines) {
      if (cachealigned_malloc(info_cachelines * kCacheline, &h->info_arena))
        throw std::runtime_error("out of memory");
    }
    uint32_t max_vrec = 0;
    uintptr_t reader_cachelines = 0;
    if (PgfiInitPhase2(ctrl, 1, 1, 0, 0, h->info.raw_variant_ct, &max_vrec, &h->info,
                       h->info_arena, &reader_cachelines, errbuf)) {
      throw std::runtime_error(std::string("PgfiInitPhase2: ") + errbuf);
    }
    PreinitPgr(&h->reader);
    const uint32_t n = h->info.raw_sample_ct;
    const uintptr_t main_bytes = reader_cachelines * kCacheline;
    const uintptr_t subset_bytes = DivUp(n, kBitsPerVec) * kBytesPerVec;
    const uintptr_t cumpop_bytes = DivUp(n, kBitsPerWord * kInt32PerVec) * kBytesPerVec;
    const uintptr_t geno_bytes = DivUp(n, kNypsPerVec) * kBytesPerVec;
    const uintptr_t dosage_bytes = DivUp(n, 2 * kInt32PerVec) * kBytesPerVec;
    // same (generous) arena size as the reference provider, loader.cpp:338-343
    const uintptr_t total = main_bytes + (2 * kPglNypTransposeBatch + 5) * subset_bytes +
                            cumpop_bytes + (1 + kPglNypTransposeBatch) * geno_bytes +
                            dosage_bytes + kPglBitTransposeBufbytes + 64;
    if (cachealigned_malloc(total, &h->reader_arena)) throw std::runtime_error("out of memory");
    PglErr rc = PgrInit(path, max_vrec, &h->info, &h->reader, h->reader_arena);
    if (rc != kPglRetSuccess) throw std::runtime_error("PgrInit failed");
    h->reader_live = true;
    unsigned char* p = h->reader_arena + main_bytes;
    h->include_vec = reinterpret_cast<uintptr_t*>(p);
    p += 2 * subset_bytes;
    p += cumpop_bytes;
    h->genovec = reinterpret_cast<uintptr_t*>(p);
    p += geno_bytes;
    h->dosage_present = reinterpret_cast<uintptr_t*>(p);
    p += subset_bytes;
    h->dosage_main = reinterpret_cast<uint16_t*>(p);
    memset(&h->pssi, 0, sizeof(h->pssi));
    return h;
  } catch (const std::exception& e) {
    g_err = e.what();
    ref_pgen_close(h);
    return nullptr;
  }
}

int ref_pgen_variant_ct(void* hv) { return int(static_cast<RefPgen*>(hv)->info.raw_variant_ct); }
int ref_pgen_sample_ct(void* hv) { return int(static_cast<RefPgen*>(hv)->info.raw_sample_ct); }

// Raw PgrGet1D output: 2-bit genovec (ceil(N/4) bytes), dosage_present bitmap
// (ceil(N/8) bytes), dosage_main (dosage_ct u16).  Returns dosage_ct, <0 on error.
int ref_pgen_get1d(void* hv, int index, unsigned char* geno_out, unsigned char* present_out,
                   uint16_t* dosage_out) {
  using namespace plink2;
  RefPgen* h = static_cast<RefPgen*>(hv);
  try {
    const uint32_t n = h->info.raw_sample_ct;
    uint32_t dosage_ct = 0;
    PglErr rc = PgrGet1D(h->include_vec, h->pssi, n, uint32_t(index), 1, &h->reader, h->genovec,
                         h->dosage_present, h->dosage_main, &dosage_ct);
    if (rc != kPglRetSuccess) throw std::runtime_error("PgrGet1D error " + std::to_string(int(rc)));
    if (geno_out) memcpy(geno_out, h->genovec, (n + 3) / 4);
    if (dosage_ct) {
      if (present_out) memcpy(present_out, h->dosage_present, (n + 7) / 8);
      if (dosage_out) memcpy(dosage_out, h->dosage_main, size_t(dosage_ct) * 2);
    }
    return int(dosage_ct);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// One SNP column as the reference provider delivers it to OpenMx:
// out[destRows] doubles (NA = R NA_real_), *maf as in loader.cpp:386-394.
// row_filter: srcRows ints, nonzero = skip that person (LoadDataAPI.h:80); may be NULL.
int ref_pgen_load_row(void* hv, int index, const int* row_filter, double* out, double* maf) {
  using namespace plink2;
  RefPgen* h = static_cast<RefPgen*>(hv);
  try {
    const uint32_t n = h->info.raw_sample_ct;
    if (index < 0 || uint32_t(index) >= h->info.raw_variant_ct) {
      throw std::runtime_error("out of data (record " + std::to_string(index + 1) +
                               " requested but only " +
                               std::to_string(h->info.raw_variant_ct) + " in file)");
    }
    uint32_t dosage_ct = 0;
    PglErr rc = PgrGet1D(h->include_vec, h->pssi, n, uint32_t(index), 1, &h->reader, h->genovec,
                         h->dosage_present, h->dosage_main, &dosage_ct);
    if (rc != kPglRetSuccess) throw std::runtime_error("PgrGet1D error " + std::to_string(int(rc)));
    const double lut[4] = {0.0, 1.0, 2.0, r_na_real()};
    std::vector<int> dest_of(n);
    int dest_rows = 0;
    for (uint32_t s = 0; s < n; ++s) {
      const bool skip = row_filter && row_filter[s];
      dest_of[s] = skip ? -1 : dest_rows++;
      if (skip) continue;
      const uintptr_t w = h->genovec[s / kBitsPerWordD2];
      out[dest_of[s]] = lut[(w >> (2 * (s % kBitsPerWordD2))) & 3];
    }
    if (dosage_ct) {
      uint32_t k = 0;
      for (uint32_t s = 0; s < n && k < dosage_ct; ++s) {
        if ((h->dosage_present[s / kBitsPerWord] >> (s % kBitsPerWord)) & 1) {
          if (dest_of[s] >= 0) out[dest_of[s]] = double(h->dosage_main[k]) * 0.00006103515625;
          ++k;
        }
      }
    }
    double total = 0;
    int finite_ct = 0;
    for (int r = 0; r < dest_rows; ++r) {
      if (!std::isfinite(out[r])) continue;
      total += out[r];
      ++finite_ct;
    }
    double m = total / (2.0 * finite_ct);
    if (m > 0.5) m = 1 - m;
    if (maf) *maf = m;
    return dest_rows;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

void ref_pgen_close(void* hv) {
  using namespace plink2;
  RefPgen* h = static_cast<RefPgen*>(hv);
  if (!h) return;
  PglErr err = kPglRetSuccess;
  if (h->reader_live) CleanupPgr(&h->reader, &err);
  if (h->info_live) CleanupPgfi(&h->info, &err);
  if (h->reader_arena) aligned_free(h->reader_arena);
  if (h->info_arena) aligned_free(h->info_arena);
  delete h;
}

}  // extern "C"
