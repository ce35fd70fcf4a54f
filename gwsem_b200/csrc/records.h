// Launch interfaces of records.cu (device-side record decoding).
#pragma once
#include "common.h"

namespace gwsem {

struct RecordBatch {
  const uint8_t* bytes = nullptr;     // concatenated record bytes (device), 16 spare bytes at the end
  const uint64_t* off = nullptr;      // n + 1 offsets into bytes
  const uint8_t* vrtype = nullptr;    // n
  const int32_t* base_row = nullptr;  // n: output row holding the LD base of record r, or -1
  int n = 0;
  int n_samples = 0;
  int n_pad = 0;
  int plink1 = 0;
  uint8_t* rows = nullptr;            // out: n rows of `pitch` bytes, 2-bit PLINK 2 coding
  int64_t pitch = 0;
  uint16_t* dosage = nullptr;         // out: n * n_pad (0xFFFF = no dosage), or nullptr when no record has one
  int32_t* err = nullptr;             // out: n (0 ok, else the pgenlib error code the reference would report)
};

struct BgenBatch {
  const uint8_t* bytes = nullptr;     // concatenated inflated probability blocks (device)
  const uint64_t* off = nullptr;      // n + 1
  int n = 0;
  int n_samples = 0;
  int layout = 1;
  const int32_t* dest_of = nullptr;   // rowFilter compaction, or nullptr for file order
  const unsigned long long* rowmask = nullptr;  // people counted in MAF when dest_of == nullptr
  int64_t out_stride = 0;
  double* out = nullptr;              // n * out_stride
  double* maf = nullptr;              // n
  int32_t* err = nullptr;             // n
};

void launch_pgen_records(gwsem_engine* e, const RecordBatch& rb, bool any_ld);
void launch_rows_to_doubles(gwsem_engine* e, const uint8_t* d_rows, int64_t pitch, const uint16_t* d_dosage, int n_pad,
                            int n_snps, int n_samples, const int32_t* d_dest_of, const unsigned long long* d_rowmask,
                            int64_t out_stride, double* d_out, double* d_maf);
void launch_bgen_blocks(gwsem_engine* e, const BgenBatch& bb);

}  // namespace gwsem
