// Host-side container parsing for the genotype sources.  What the reference does in
// PgfiInitPhase1/2 (src/pgenlib_read.cc:681-977, 999-1677) and in bgen View / IndexQuery
// (src/bgen.cpp:73-115, 506-664, 714-934), reduced to what the GPU path needs: where each
// variant's record lives, how it is encoded, and (bgen) the index order.
#include "source.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstring>

#include "common.h"

using gwsem::fail;

gwsem_source::~gwsem_source() {
  if (fd >= 0) close(fd);
}

void gwsem_source::read_at(uint64_t off, void* dst, size_t len) const {
  uint8_t* p = static_cast<uint8_t*>(dst);
  while (len) {
    ssize_t got = pread(fd, p, len, (off_t)off);
    if (got <= 0) fail("%s: read failure at offset %llu", path.c_str(), (unsigned long long)off);
    p += got;
    off += (uint64_t)got;
    len -= (size_t)got;
  }
}

uint32_t gwsem_source::ld_base(uint32_t v) const {
  if (mode != 0x10) return v;
  uint32_t b = v;
  while ((vrtype[b] & 6) == 2) {  // main-track types 2 and 3 are LD-compressed
    if (b == 0) fail("%s: LD-compressed record %u has no base", path.c_str(), v);
    --b;
  }
  return b;
}

namespace {

bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

void open_pgen(gwsem_source* s, int src_rows) {
  uint8_t hdr[12] = {0};
  if (s->file_size < 3) fail("%s is too small to be a valid .pgen file", s->path.c_str());
  s->read_at(0, hdr, s->file_size < 12 ? 3 : 12);
  if (hdr[0] != 0x6c || hdr[1] != 0x1b)
    fail("%s is not a .pgen file (first two bytes don't match the magic number)", s->path.c_str());
  s->mode = hdr[2];
  if (s->mode == 0x00) fail("%s: sample-major PLINK 1 .bed files are not supported", s->path.c_str());
  if (s->mode == 0x01) {
    s->fixed_vrtype = 0;
    s->fixed_first = 3;
    if (src_rows <= 0) {
      // numAvailableRecords on a .bed without sample information: the reference reports NA
      // (loader.cpp:295-298, tests/testthat/test-formats.R:25-33)
      s->variants_known = false;
      return;
    }
    s->n_samples = (uint32_t)src_rows;
    s->fixed_width = (s->n_samples + 3) / 4;
    if ((s->file_size - 3) % s->fixed_width)
      fail("%s: unexpected PLINK 1 .bed file size (since raw_sample_ct was %u, [file size - 3] should be divisible by %llu)",
           s->path.c_str(), s->n_samples, (unsigned long long)s->fixed_width);
    s->n_variants = (uint32_t)((s->file_size - 3) / s->fixed_width);
    return;
  }
  if (s->file_size < 12) fail("%s is too small to be a valid .pgen file", s->path.c_str());
  memcpy(&s->n_variants, hdr + 3, 4);
  memcpy(&s->n_samples, hdr + 7, 4);
  const uint32_t ctrl = hdr[11];
  if (src_rows > 0 && (uint32_t)src_rows != s->n_samples)
    fail("PgfiInitPhase1() was called with raw_sample_ct == %d, but %s contains %u samples", src_rows, s->path.c_str(),
         s->n_samples);
  if (s->n_samples == 0) fail("pgen file '%s' has no samples", s->path.c_str());
  uint64_t pos = 12;
  if (s->mode >= 0x02 && s->mode <= 0x04) {
    if (ctrl & 63) fail("%s: fixed-width storage mode with a variable-width header byte", s->path.c_str());
    if ((ctrl >> 6) == 3) pos += (s->n_variants + 7) / 8;
    s->fixed_width = (s->n_samples + 3) / 4;
    if (s->mode == 0x03) { s->fixed_width += 2ull * s->n_samples; s->fixed_vrtype = 0x40; }
    if (s->mode == 0x04) { s->fixed_width += 4ull * s->n_samples; s->fixed_vrtype = 0xc0; }
    s->any_dosage = s->mode != 0x02;
    s->fixed_first = pos;
    if (pos + s->fixed_width * s->n_variants != s->file_size)
      fail("%s: unexpected .pgen file size (expected %llu bytes)", s->path.c_str(),
           (unsigned long long)(pos + s->fixed_width * s->n_variants));
    return;
  }
  if (s->mode != 0x10)
    fail("%s: third byte does not correspond to a storage mode supported by this reader", s->path.c_str());
  if (ctrl & 8) fail("%s: fused vrtype/record-length headers (single-sample files) are not supported", s->path.c_str());
  // A header with stored ALT-allele counts is a multiallelic file.  The reference provider opens it with
  // allele_idx_offsets == NULL (src/loader.cpp:323-327), which pgenlib rejects (src/pgenlib_read.cc:1172-1178); the
  // aux1 tracks of such records (SkipAux1, src/pgenlib_read.cc:6605-6620) are therefore never reached from gwsem.
  if ((ctrl >> 4) & 3)
    fail("%s: PgfiInitPhase2: pgfip->allele_idx_offsets must be allocated before PgfiInitPhase2() is called "
         "(multiallelic .pgen: not readable through gwsem's provider)", s->path.c_str());
  const bool nonref_stored = (ctrl >> 6) == 3;
  const bool wide = (ctrl & 4) != 0;
  const uint32_t len_bytes = 1 + (ctrl & 3);
  const uint32_t M = s->n_variants;
  const uint32_t vblocks = (M + 65535) / 65536;
  uint64_t cur = 0;
  s->read_at(pos, &cur, 8);
  pos += 8ull * vblocks;
  s->vrtype.assign(M + 1, 0);
  s->fpos.assign(M + 1, 0);
  std::vector<uint8_t> buf;
  uint32_t v = 0;
  for (uint32_t b = 0; b < vblocks; ++b) {
    const uint32_t cnt = (b + 1 == vblocks) ? M - 65536u * b : 65536u;
    const uint64_t type_bytes = wide ? cnt : (cnt + 1) / 2;
    const uint64_t block_bytes = type_bytes + (uint64_t)cnt * len_bytes + (nonref_stored ? (cnt + 7) / 8 : 0);
    if (pos + block_bytes > s->file_size) fail("%s: header runs past the end of the file", s->path.c_str());
    buf.resize(block_bytes + 8);
    s->read_at(pos, buf.data(), block_bytes);
    for (uint32_t k = 0; k < cnt; ++k)
      s->vrtype[v + k] = wide ? buf[k] : (uint8_t)((buf[k / 2] >> (4 * (k & 1))) & 15);
    const uint8_t* lens = buf.data() + type_bytes;
    for (uint32_t k = 0; k < cnt; ++k) {
      uint32_t rl = 0;
      memcpy(&rl, lens + (uint64_t)k * len_bytes, len_bytes);
      s->fpos[v + k] = cur;
      cur += rl;
      if (s->vrtype[v + k] & 0x60) s->any_dosage = true;
    }
    pos += block_bytes;
    v += cnt;
  }
  s->fpos[M] = cur;
  if (cur > s->file_size) fail("%s: variant records run past the end of the file", s->path.c_str());
}

// ---- sqlite (.bgi) through the runtime library: no headers are installed in this image
struct Sqlite {
  void* lib = nullptr;
  int (*open_v2)(const char*, void**, int, const char*) = nullptr;
  int (*prepare_v2)(void*, const char*, int, void**, const char**) = nullptr;
  int (*step)(void*) = nullptr;
  long long (*column_int64)(void*, int) = nullptr;
  int (*finalize)(void*) = nullptr;
  int (*close_db)(void*) = nullptr;
  bool load() {
    for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      lib = dlopen(name, RTLD_NOW);
      if (lib) break;
    }
    if (!lib) return false;
    open_v2 = (decltype(open_v2))dlsym(lib, "sqlite3_open_v2");
    prepare_v2 = (decltype(prepare_v2))dlsym(lib, "sqlite3_prepare_v2");
    step = (decltype(step))dlsym(lib, "sqlite3_step");
    column_int64 = (decltype(column_int64))dlsym(lib, "sqlite3_column_int64");
    finalize = (decltype(finalize))dlsym(lib, "sqlite3_finalize");
    close_db = (decltype(close_db))dlsym(lib, "sqlite3_close");
    return open_v2 && prepare_v2 && step && column_int64 && finalize && close_db;
  }
};

void open_bgen(gwsem_source* s) {
  uint8_t head[24];
  if (s->file_size < 24) fail("%s is too small to be a valid .bgen file", s->path.c_str());
  s->read_at(0, head, 24);
  uint32_t offset, hdr_len, M, N;
  memcpy(&offset, head, 4);
  memcpy(&hdr_len, head + 4, 4);
  memcpy(&M, head + 8, 4);
  memcpy(&N, head + 12, 4);
  if (memcmp(head + 16, "bgen", 4) != 0 && memcmp(head + 16, "\0\0\0\0", 4) != 0)
    fail("%s: not a bgen file (bad magic)", s->path.c_str());
  if (hdr_len < 20) fail("%s: malformed bgen header", s->path.c_str());
  s->read_at(4 + (uint64_t)hdr_len - 4, &s->bgen_flags, 4);
  s->bgen_compression = s->bgen_flags & 3;
  s->bgen_layout = (s->bgen_flags >> 2) & 15;
  if (s->bgen_layout != 1 && s->bgen_layout != 2) fail("%s: bgen layout %u is not supported", s->path.c_str(), s->bgen_layout);
  if (s->bgen_compression > 2) fail("%s: unknown compression", s->path.c_str());
  s->n_samples = N;
  // variant order = the .bgi index order (src/bgen.cpp:649-663)
  const std::string bgi = s->path + ".bgi";
  struct stat st;
  if (stat(bgi.c_str(), &st) != 0) fail("%s: index file not found (generate it with bgenix)", bgi.c_str());
  Sqlite q;
  if (!q.load()) fail("cannot load the sqlite3 runtime library needed to read %s", bgi.c_str());
  void* db = nullptr;
  if (q.open_v2(bgi.c_str(), &db, 1 /* SQLITE_OPEN_READONLY */, nullptr) != 0) fail("cannot open %s", bgi.c_str());
  void* stmt = nullptr;
  const char* sql =
      "SELECT file_start_position, size_in_bytes FROM Variant ORDER BY chromosome, position, rsid, allele1, allele2";
  if (q.prepare_v2(db, sql, -1, &stmt, nullptr) != 0) {
    q.close_db(db);
    fail("%s: not a bgenix index (no Variant table)", bgi.c_str());
  }
  while (q.step(stmt) == 100 /* SQLITE_ROW */) {
    gwsem_source::BgenVariant v;
    v.start = (uint64_t)q.column_int64(stmt, 0);
    v.next = v.start + (uint64_t)q.column_int64(stmt, 1);
    if (v.next > s->file_size) {
      q.finalize(stmt);
      q.close_db(db);
      fail("%s: index points past the end of %s", bgi.c_str(), s->path.c_str());
    }
    s->bgen.push_back(v);
  }
  q.finalize(stmt);
  q.close_db(db);
  s->n_variants = (uint32_t)s->bgen.size();
  (void)M;
  (void)offset;
}

}  // namespace

void gwsem_source::parse_bgen_variant(uint32_t index) {
  BgenVariant& v = bgen[index];
  if (v.parsed) return;
  const uint64_t span = std::min<uint64_t>(v.next - v.start, 1 << 20);
  std::vector<uint8_t> buf(span + 8);
  read_at(v.start, buf.data(), span);
  const uint8_t* p = buf.data();
  const uint8_t* end = p + span;
  auto take = [&](uint32_t width, std::string* out) {
    uint32_t n = 0;
    if (p + width > end) fail("%s: truncated variant block", path.c_str());
    memcpy(&n, p, width);
    p += width;
    if (p + n > end) fail("%s: truncated variant block", path.c_str());
    out->assign(reinterpret_cast<const char*>(p), n);
    p += n;
  };
  if (bgen_layout == 1) {
    uint32_t n = 0;
    memcpy(&n, p, 4);
    p += 4;
    if (n != n_samples) fail("%s: variant block sample count %u != %u", path.c_str(), n, n_samples);
  }
  take(2, &v.snpid);
  take(2, &v.rsid);
  take(2, &v.chrom);
  memcpy(&v.pos, p, 4);
  p += 4;
  uint32_t K = 2;
  if (bgen_layout == 2) {
    K = 0;
    memcpy(&K, p, 2);
    p += 2;
  }
  if (K != 2) fail("%s: variant %s has %u alleles; only biallelic variants are supported", path.c_str(), v.rsid.c_str(), K);
  take(4, &v.a1);
  take(4, &v.a2);
  v.geno = v.start + (uint64_t)(p - buf.data());
  v.parsed = true;
}

namespace gwsem {

gwsem_source* open_source(const char* path, const char* method, int src_rows) {
  auto* s = new gwsem_source;
  try {
    s->path = path;
    const std::string m = method ? method : "";
    if (m == "pgen") s->kind = gwsem_source::kPgen;
    else if (m == "bgen") s->kind = gwsem_source::kBgen;
    else if (m.empty()) s->kind = ends_with(s->path, ".bgen") ? gwsem_source::kBgen : gwsem_source::kPgen;
    else fail("unknown data provider method '%s'", method);
    s->fd = open(path, O_RDONLY);
    if (s->fd < 0) fail("Failed to open %s : %s", path, strerror(errno));
    struct stat st;
    if (fstat(s->fd, &st) != 0) fail("Failed to stat %s", path);
    s->file_size = (uint64_t)st.st_size;
    if (s->kind == gwsem_source::kPgen) {
      open_pgen(s, src_rows);
      s->checkpoint_columns = "MAF";  // loader.cpp:188-192
    } else {
      open_bgen(s);
      if (src_rows > 0 && (uint32_t)src_rows != s->n_samples)
        fail("data has %d rows but %s has %d samples", src_rows, path, (int)s->n_samples);  // loader.cpp:106-110
      s->checkpoint_columns = "SNP\tRSID\tCHR\tBP\tA2\tA1\tMAF";  // loader.cpp:70-80
    }
    s->load_counter = 1;  // the file is opened exactly once (tests/testthat/test-formats.R:48)
    return s;
  } catch (...) {
    delete s;
    throw;
  }
}

void bgen_inflate(const gwsem_source& s, const uint8_t* payload, uint32_t payload_len, std::vector<uint8_t>& dst) {
  if (s.bgen_compression == 0) {
    if (payload_len != dst.size()) fail("%s: stored block has %u bytes, expected %zu", s.path.c_str(), payload_len, dst.size());
    memcpy(dst.data(), payload, payload_len);
    return;
  }
  if (s.bgen_compression == 1) {
    uLongf got = (uLongf)dst.size();
    if (uncompress(dst.data(), &got, payload, payload_len) != Z_OK || got != dst.size())
      fail("%s: zlib inflate of a genotype block failed", s.path.c_str());
    return;
  }
  typedef size_t (*zstd_fn)(void*, size_t, const void*, size_t);
  static zstd_fn fn = nullptr;
  if (!fn) {
    void* lib = dlopen("libzstd.so.1", RTLD_NOW);
    if (lib) fn = (zstd_fn)dlsym(lib, "ZSTD_decompress");
    if (!fn) fail("%s is zstd-compressed and the zstd runtime library is not available", s.path.c_str());
  }
  if (fn(dst.data(), dst.size(), payload, payload_len) != dst.size())
    fail("%s: zstd inflate of a genotype block failed", s.path.c_str());
}

}  // namespace gwsem
