// Genotype file sources: host-side container parsing for .pgen / .bed / .bgen.  Records are
// handed to the GPU still encoded; all per-genotype decoding happens in the kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

struct gwsem_source {
  enum Kind { kPgen, kBgen } kind = kPgen;
  std::string path;
  int fd = -1;
  uint64_t file_size = 0;
  uint32_t n_variants = 0, n_samples = 0;
  bool variants_known = true;
  int load_counter = 0;
  std::string checkpoint_columns;

  // ---- pgen / bed (format: reference src/pgenlib_misc.h:614-905)
  int mode = 0;                    // 0x01 bed, 0x02..0x04 fixed width, 0x10 variable
  uint64_t fixed_first = 0;        // offset of record 0 (fixed-width modes)
  uint64_t fixed_width = 0;
  uint8_t fixed_vrtype = 0;
  std::vector<uint8_t> vrtype;     // variable mode
  std::vector<uint64_t> fpos;      // variable mode, n_variants + 1
  bool any_dosage = false;

  // ---- bgen (reference src/bgen.cpp:73-115, src/include/genfile/bgen/bgen.hpp:479-541)
  uint32_t bgen_flags = 0, bgen_layout = 0, bgen_compression = 0;
  struct BgenVariant {
    uint64_t start = 0, geno = 0, next = 0;  // identifying block, genotype block, one past the end
    bool parsed = false;
    std::string snpid, rsid, chrom, a1, a2;
    uint32_t pos = 0;
  };
  std::vector<BgenVariant> bgen;   // index (.bgi) order

  ~gwsem_source();
  void read_at(uint64_t off, void* dst, size_t len) const;

  uint8_t record_vrtype(uint32_t v) const { return mode == 0x10 ? vrtype[v] : fixed_vrtype; }
  uint64_t record_offset(uint32_t v) const { return mode == 0x10 ? fpos[v] : fixed_first + fixed_width * v; }
  uint64_t record_length(uint32_t v) const { return mode == 0x10 ? fpos[v + 1] - fpos[v] : fixed_width; }
  // most recent non-LD variant at or before v (pgenlib_read.cc:1679-1726)
  uint32_t ld_base(uint32_t v) const;
  void parse_bgen_variant(uint32_t index);
};

namespace gwsem {
gwsem_source* open_source(const char* path, const char* method, int src_rows);
// inflate one bgen genotype block (zlib / zstd / stored) into dst (already sized)
void bgen_inflate(const gwsem_source& s, const uint8_t* payload, uint32_t payload_len, std::vector<uint8_t>& dst);
}  // namespace gwsem
