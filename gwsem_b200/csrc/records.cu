// Device-side decoding of encoded genotype records into the engine's HBM layout:
//   * PGEN variant records of every main-track type (reference src/pgenlib_misc.h:720-760,
//     src/pgenlib_read.cc:2428-2535, 2789-2840), difflists (pgenlib_misc.h:699-718,
//     pgenlib_read.cc:2008-2042, 2325-2372), hardcall-phase skipping (6706-6725) and the three
//     dosage track layouts (7065-7318)  ->  2-bit rows (PLINK 2 coding) + a u16 dosage plane
//   * BGEN probability blocks, layout 1 and layout 2 (src/include/genfile/bgen/bgen.hpp:783-811,
//     1025-1052, 815-885, 1139-1224) -> the dosage column of BgenXfer (src/loader.cpp:38-50)
//   * rows (+ dosage plane) -> double columns with rowFilter compaction and MAF
//     (src/loader.cpp:211-273, 386-394)
// One CTA per record; the serial parts of the formats (varint chains) are cut at the 64-entry
// group boundaries the format provides, one thread per group.
#include "common.h"
#include "kernels.h"
#include "records.h"

namespace gwsem {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxGroups = 4096;  // 64-entry difflist groups per record (covers 2M samples)

__device__ __forceinline__ double na_real() { return __longlong_as_double((long long)GWSEM_NA_BITS); }

__device__ __forceinline__ uint32_t ld16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t ld_n(const uint8_t* p, int n) {
  uint32_t v = 0;
  for (int k = 0; k < n; ++k) v |= (uint32_t)p[k] << (8 * k);
  return v;
}
// up to 4 bytes, never touching p[k] for k >= avail
__device__ __forceinline__ uint32_t ld32_guard(const uint8_t* p, int64_t avail) {
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < avail) v |= (uint32_t)p[k] << (8 * k);
  return v;
}
__device__ __forceinline__ uint32_t varint(const uint8_t*& p, const uint8_t* end) {
  uint32_t v = 0, shift = 0;
  while (p < end) {
    const uint32_t b = *p++;
    v |= (b & 127u) << shift;
    if (b <= 127u) return v;
    shift += 7;
  }
  return 0x80000000u;
}
__device__ __forceinline__ uint32_t spread16(uint32_t x) {  // bit k -> bit 2k
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ uint32_t bed_to_plink2_w(uint32_t w) {
  const uint32_t hi = w & 0xAAAAAAAAu, lo = w & 0x55555555u;
  return (~hi & 0xAAAAAAAAu) | ((hi >> 1) ^ lo);
}

struct ListHeader {
  uint32_t len, groups, id_bytes;
  const uint8_t* group_ids;
  const uint8_t* group_lens;
  const uint8_t* values;   // packed 2-bit genotypes (difflist only)
  const uint8_t* deltas;   // start of group 0's varints
  const uint8_t* end;      // one past the list (valid after list_find_end)
  int error;
};

// Thread 0 parses the fixed part of a difflist / deltalist and the per-group start offsets.
__device__ void list_prepare(const uint8_t* p, const uint8_t* rec_end, uint32_t n_samples, bool with_values,
                             ListHeader* h, uint32_t* gstart) {
  h->error = 0;
  h->len = varint(p, rec_end);
  h->groups = 0;
  h->end = p;
  if (h->len == 0) return;
  if (h->len > n_samples / 8) { h->error = 1; h->len = 0; return; }
  const uint32_t groups = (h->len + 63) / 64;
  if (groups > kMaxGroups) { h->error = 2; h->len = 0; return; }
  uint32_t ib = 1;
  while (ib < 4 && (n_samples >> (8 * ib))) ++ib;
  h->groups = groups;
  h->id_bytes = ib;
  h->group_ids = p;
  p += (uint64_t)groups * ib;
  h->group_lens = p;
  p += groups - 1;
  h->values = p;
  if (with_values) p += (h->len + 3) / 4;
  h->deltas = p;
  if (p > rec_end) { h->error = 1; h->len = 0; h->groups = 0; return; }
  uint32_t acc = 0;
  for (uint32_t g = 0; g < groups; ++g) {
    gstart[g] = acc;
    if (g + 1 < groups) acc += (uint32_t)h->group_lens[g] + 63u;
  }
}

// Walk group g; fn(entry index, sample id).  Returns the position after the group's varints.
template <typename Fn>
__device__ const uint8_t* list_walk_group(const ListHeader& h, const uint32_t* gstart, uint32_t g,
                                          const uint8_t* rec_end, uint32_t n_samples, int* err, Fn fn) {
  const uint8_t* p = h.deltas + gstart[g];
  uint32_t sid = ld_n(h.group_ids + (uint64_t)g * h.id_bytes, (int)h.id_bytes);
  const uint32_t first = g * 64, last = min(h.len, first + 64);
  for (uint32_t k = first; k < last; ++k) {
    if (k != first) sid += varint(p, rec_end);
    if (sid >= n_samples) { *err = 1; return p; }
    fn(k, sid);
  }
  return p;
}

__device__ __forceinline__ void set_code(uint32_t* row, uint32_t sid, uint32_t code) {
  const uint32_t sh = 2 * (sid & 15);
  atomicAnd(&row[sid >> 4], ~(3u << sh));
  if (code) atomicOr(&row[sid >> 4], code << sh);
}

// pass 0: records whose main track stands alone; pass 1: LD-compressed records (their base row,
// decoded in pass 0, is base_row[r] of this batch's output).
__global__ void __launch_bounds__(kThreads)
    pgen_records_kernel(RecordBatch rb, int pass) {
  __shared__ ListHeader lh;
  __shared__ uint32_t gstart[kMaxGroups];
  __shared__ uint32_t scan[kThreads];
  __shared__ const uint8_t* cursor;
  __shared__ int err_sh;
  __shared__ uint32_t red;

  const int r = blockIdx.x;
  const uint32_t vt = rb.plink1 ? 0u : rb.vrtype[r];
  const uint32_t mt = vt & 7u;
  const bool is_ld = !rb.plink1 && (mt == 2u || mt == 3u);
  if ((int)is_ld != pass) return;
  const int tid = threadIdx.x;
  const uint32_t N = (uint32_t)rb.n_samples;
  const uint8_t* p = rb.bytes + rb.off[r];
  const uint8_t* rec_end = rb.bytes + rb.off[r + 1];
  uint32_t* row = reinterpret_cast<uint32_t*>(rb.rows + (int64_t)r * rb.pitch);
  const int row_words = (int)(rb.pitch / 4);
  const int used_words = (int)((N + 15) / 16);
  if (tid == 0) { err_sh = 0; cursor = p; }
  __syncthreads();

  // ---- main track: base fill
  if (rb.plink1 || mt == 0u) {
    const int64_t nbytes = (N + 3) / 4;
    for (int w = tid; w < row_words; w += kThreads) {
      uint32_t v = 0;
      if (w < used_words) {
        v = ld32_guard(p + 4 * (int64_t)w, min((int64_t)(nbytes - 4 * (int64_t)w), (int64_t)(rec_end - (p + 4 * (int64_t)w))));
        if (rb.plink1) v = bed_to_plink2_w(v);
      }
      row[w] = v;
    }
    if (tid == 0) cursor = p + nbytes;
  } else if (mt == 1u) {
    const uint32_t c2 = p[0];
    const uint32_t lo = (c2 >> 2) & 3u, delta = c2 & 3u;
    const uint8_t* bits = p + 1;
    const int64_t nbytes = (N + 7) / 8;
    for (int w = tid; w < row_words; w += kThreads) {
      uint32_t v = 0;
      if (w < used_words) {
        const int64_t avail = nbytes - 2 * (int64_t)w;
        const uint32_t b16 = (uint32_t)bits[2 * (int64_t)w] | (avail > 1 ? (uint32_t)bits[2 * (int64_t)w + 1] << 8 : 0u);
        v = lo * 0x55555555u + spread16(b16) * delta;
      }
      row[w] = v;
    }
    if (tid == 0) cursor = p + 1 + nbytes;
  } else if (is_ld) {
    const uint32_t* base = reinterpret_cast<const uint32_t*>(rb.rows + (int64_t)rb.base_row[r] * rb.pitch);
    for (int w = tid; w < row_words; w += kThreads) row[w] = base[w];
  } else {  // 4, 6, 7: constant; 5: all reference
    const uint32_t fill = (mt == 5u ? 0u : (mt & 3u)) * 0x55555555u;
    for (int w = tid; w < row_words; w += kThreads) row[w] = w < used_words ? fill : 0u;
  }
  __syncthreads();

  // ---- main track: difflist patch (types 1, 2, 3, 4, 6, 7)
  if (!rb.plink1 && mt != 0u && mt != 5u) {
    if (tid == 0) list_prepare(cursor, rec_end, N, true, &lh, gstart);
    __syncthreads();
    if (lh.error && tid == 0) err_sh = 6;
    for (uint32_t g = tid; g < lh.groups; g += kThreads) {
      int e = 0;
      const uint8_t* after = list_walk_group(lh, gstart, g, rec_end, N, &e, [&](uint32_t k, uint32_t sid) {
        set_code(row, sid, (lh.values[k >> 2] >> (2 * (k & 3))) & 3u);
      });
      if (e) err_sh = 6;
      if (g + 1 == lh.groups) cursor = after;
    }
    if (tid == 0 && lh.groups == 0) cursor = lh.end;
    __syncthreads();
    if (mt == 3u) {  // inverted LD: 0 <-> 2 after patching (pgenlib_read.cc:1994-2006)
      for (int w = tid; w < used_words; w += kThreads) row[w] ^= (~row[w] & 0x55555555u) << 1;
      __syncthreads();
    }
  }
  if ((N & 15u) && tid == 0) row[used_words - 1] &= (1u << (2 * (N & 15u))) - 1u;  // zero trailing entries
  __syncthreads();

  // ---- tracks after the main one only matter when a dosage track is present (IMPLPgrGetD)
  uint16_t* dos = rb.dosage ? rb.dosage + (int64_t)r * rb.n_pad : nullptr;
  if (dos)
    for (int s = tid; s < rb.n_pad; s += kThreads) dos[s] = 0xFFFF;
  if (rb.plink1 || !(vt & 0x60u) || !dos) {
    if (tid == 0) rb.err[r] = err_sh;
    return;
  }
  if (vt & 8u) {  // multiallelic hardcalls: the reference provider cannot read these either
    if (tid == 0) rb.err[r] = 8;
    return;
  }
  if (vt & 0x10u) {  // skip the hardcall-phase track: needs the het count
    uint32_t het = 0;
    for (int w = tid; w < used_words; w += kThreads) het += __popc(row[w] & ~(row[w] >> 1) & 0x55555555u);
    if (tid == 0) red = 0;
    __syncthreads();
    atomicAdd(&red, het);
    __syncthreads();
    het = red;
    __syncthreads();
    const uint8_t* q = cursor;
    const uint32_t first = 1 + het / 8;
    if (het == 0) {
      if (tid == 0) err_sh = 6;
    } else if (q[0] & 1) {
      uint32_t pc = 0;
      for (uint32_t k = tid; k < first; k += kThreads) pc += __popc((uint32_t)q[k]);
      if (tid == 0) red = 0;
      __syncthreads();
      atomicAdd(&red, pc);
      __syncthreads();
      if (tid == 0) {
        if (red < 2) err_sh = 6;
        else cursor = q + first + (red - 1 + 7) / 8;
      }
    } else if (tid == 0) {
      cursor = q + first;
    }
    __syncthreads();
  }
  if (err_sh) {
    if (tid == 0) rb.err[r] = err_sh;
    return;
  }
  const uint8_t* q = cursor;
  const uint32_t dt = vt & 0x60u;
  if (dt == 0x40u) {  // unconditional: N values, 65535 = missing = no entry
    for (uint32_t s = tid; s < N; s += kThreads) dos[s] = (uint16_t)ld16(q + 2ull * s);
  } else if (dt == 0x60u) {  // presence bitarray, then one value per set bit
    const uint8_t* vals = q + (N + 7) / 8;
    const uint32_t chunks = (N + 31) / 32;
    const uint32_t per = (chunks + kThreads - 1) / kThreads;
    const uint32_t c0 = min(chunks, (uint32_t)tid * per), c1 = min(chunks, c0 + per);
    const int64_t nbytes = (N + 7) / 8;
    uint32_t total = 0;
    for (uint32_t c = c0; c < c1; ++c) total += __popc(ld32_guard(q + 4ull * c, nbytes - 4ll * c));
    scan[tid] = total;
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (int k = 0; k < kThreads; ++k) { const uint32_t t = scan[k]; scan[k] = acc; acc += t; }
    }
    __syncthreads();
    uint32_t rank = scan[tid];
    for (uint32_t c = c0; c < c1; ++c) {
      uint32_t bits = ld32_guard(q + 4ull * c, nbytes - 4ll * c);
      while (bits) {
        const uint32_t s = 32u * c + (uint32_t)__ffs((int)bits) - 1u;
        bits &= bits - 1;
        if (s < N) dos[s] = (uint16_t)ld16(vals + 2ull * rank);
        ++rank;
      }
    }
  } else {  // 0x20: delta-encoded sample id list, then one value per entry
    if (tid == 0) list_prepare(q, rec_end, N, false, &lh, gstart);
    __syncthreads();
    if (lh.error) {
      if (tid == 0) rb.err[r] = 6;
      return;
    }
    if (lh.groups) {
      if (tid == 0) {  // the values start where the last group's varints end
        int e = 0;
        cursor = list_walk_group(lh, gstart, lh.groups - 1, rec_end, N, &e, [](uint32_t, uint32_t) {});
        if (e) err_sh = 6;
      }
      __syncthreads();
      const uint8_t* vals = cursor;
      for (uint32_t g = tid; g < lh.groups; g += kThreads) {
        int e = 0;
        list_walk_group(lh, gstart, g, rec_end, N, &e,
                        [&](uint32_t k, uint32_t sid) { dos[sid] = (uint16_t)ld16(vals + 2ull * k); });
        if (e) err_sh = 6;
      }
    }
  }
  __syncthreads();
  if (tid == 0) rb.err[r] = err_sh;
}

// rows (+ dosage plane) -> doubles.  dest_of == nullptr: file order with stride out_stride (the
// statistics path); otherwise compacted to dest rows (the loadRow path).  MAF accumulators are
// integers in units of 2^-14 allele, over people with dest_of >= 0 (all people if nullptr):
// sums of dosages / hardcalls are exact, so any order reproduces the reference's double sum.
__global__ void rows_to_doubles_kernel(const uint8_t* __restrict__ rows, int64_t pitch, const uint16_t* __restrict__ dosage,
                                       int n_pad, int n_samples, const int32_t* __restrict__ dest_of,
                                       const unsigned long long* __restrict__ rowmask, int64_t out_stride,
                                       double* __restrict__ out, unsigned long long* __restrict__ maf_acc) {
  const int r = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long value = 0;
  unsigned finite = 0;
  if (i < n_samples) {
    const int d = dest_of ? dest_of[i] : i;
    const bool counted = dest_of ? d >= 0 : (rowmask ? ((rowmask[i >> 5] >> (2 * (i & 31))) & 1ull) != 0 : true);
    if (d >= 0) {
      const unsigned code = (rows[r * pitch + (i >> 2)] >> (2 * (i & 3))) & 3u;
      const unsigned ds = dosage ? dosage[(int64_t)r * n_pad + i] : 0xFFFFu;
      double v;
      unsigned units;
      if (ds != 0xFFFFu) { v = (double)ds * 0.00006103515625; units = ds; finite = 1; }
      else if (code != 3u) { v = (double)code; units = code << 14; finite = 1; }
      else { v = na_real(); units = 0; finite = 0; }
      out[(int64_t)r * out_stride + d] = v;
      if (!counted) finite = 0;
      value = finite ? units : 0;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    value += __shfl_xor_sync(0xffffffffu, value, o);
    finite += __shfl_xor_sync(0xffffffffu, finite, o);
  }
  if ((threadIdx.x & 31) == 0 && finite) {
    atomicAdd(&maf_acc[2 * r], value);
    atomicAdd(&maf_acc[2 * r + 1], (unsigned long long)finite);
  }
}

__global__ void maf_finish_units_kernel(const unsigned long long* __restrict__ acc, int n, double* __restrict__ maf) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double m = ((double)acc[2 * r] * 0.00006103515625) / (2.0 * (double)acc[2 * r + 1]);
  if (m > 0.5) m = 1 - m;
  maf[r] = m;
}

// BGEN probability block -> dosage column.  One CTA per variant.  err: 0 ok, 20 malformed,
// 21 ploidy / phased ("set_min_max_ploidy ... not implemented"), 22 missing sample ("not implemented").
__global__ void __launch_bounds__(kThreads)
    bgen_blocks_kernel(BgenBatch bb) {
  __shared__ int err_sh;
  __shared__ double sum_sh[kThreads];
  __shared__ unsigned cnt_sh[kThreads];
  const int r = blockIdx.x, tid = threadIdx.x;
  const uint32_t N = (uint32_t)bb.n_samples;
  const uint8_t* blk = bb.bytes + bb.off[r];
  const uint64_t len = bb.off[r + 1] - bb.off[r];
  if (tid == 0) err_sh = 0;
  __syncthreads();
  double sum = 0.0;
  unsigned cnt = 0;
  auto emit = [&](uint32_t s, double p0, double p1) {
    const int d = bb.dest_of ? bb.dest_of[s] : (int)s;
    if (d < 0) return;
    double dose = p1 * 1 + p0 * 2;
    if (dose == 0.0) dose = na_real();   // loader.cpp:42-43
    else if (!bb.rowmask || ((bb.rowmask[s >> 5] >> (2 * (s & 31))) & 1ull)) { sum += dose; ++cnt; }
    bb.out[(int64_t)r * bb.out_stride + d] = dose;
  };
  if (bb.layout == 1) {
    if (len != 6ull * N) { if (tid == 0) err_sh = 20; }
    else
      for (uint32_t s = tid; s < N; s += kThreads) {
        double p0 = (double)ld16(blk + 6ull * s), p1 = (double)ld16(blk + 6ull * s + 2);
        p0 /= 32768.0;
        p1 /= 32768.0;
        emit(s, p0, p1);
      }
  } else {
    bool ok = len >= 10ull + N;
    uint32_t bits = 0;
    if (ok) {
      const uint32_t n2 = ld_n(blk, 4), K = ld16(blk + 4);
      const uint32_t pmin = blk[6], pmax = blk[7], phased = blk[8 + N] & 1u;
      bits = blk[9 + N];
      if (n2 != N || K != 2 || bits < 1 || bits > 32) { ok = false; if (tid == 0) err_sh = 20; }
      else if (pmin != 2 || pmax != 2 || phased) { ok = false; if (tid == 0) err_sh = 21; }
      else if (len < 10ull + N + ((uint64_t)2 * bits * N + 7) / 8) { ok = false; if (tid == 0) err_sh = 20; }
    } else if (tid == 0) err_sh = 20;
    if (ok) {
      const uint8_t* ploidy = blk + 8;
      const uint8_t* probs = blk + 10 + N;
      const uint64_t nbytes = len - (10ull + N);
      const uint64_t mask = (~0ull) >> (64 - bits);
      const double denom = (double)mask;
      for (uint32_t s = tid; s < N; s += kThreads) {
        const int d = bb.dest_of ? bb.dest_of[s] : (int)s;
        const bool kept = bb.dest_of ? d >= 0 : (!bb.rowmask || ((bb.rowmask[s >> 5] >> (2 * (s & 31))) & 1ull));
        if (!kept) {  // set_sample() == false: the value is never looked at (bgen.hpp:1196-1204)
          if (!bb.dest_of) bb.out[(int64_t)r * bb.out_stride + s] = na_real();
          continue;
        }
        if (ploidy[s] & 0x80u) { err_sh = 22; continue; }
        uint64_t v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t bitpos = ((uint64_t)2 * s + k) * bits;
          const uint64_t byte = bitpos >> 3;
          uint64_t w = 0;
          for (int b = 0; b < 8; ++b)
            if (byte + b < nbytes) w |= (uint64_t)probs[byte + b] << (8 * b);
          v[k] = (w >> (bitpos & 7)) & mask;
        }
        emit(s, (double)v[0] / denom, (double)v[1] / denom);
      }
    }
  }
  sum_sh[tid] = sum;
  cnt_sh[tid] = cnt;
  __syncthreads();
  for (int o = kThreads / 2; o; o >>= 1) {
    if (tid < o) { sum_sh[tid] += sum_sh[tid + o]; cnt_sh[tid] += cnt_sh[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    bb.err[r] = err_sh;
    double m = sum_sh[0] / (2.0 * (double)cnt_sh[0]);   // loader.cpp:141-143
    if (m > .5) m = 1 - m;
    bb.maf[r] = m;
  }
}

}  // namespace

void launch_pgen_records(gwsem_engine* e, const RecordBatch& rb, bool any_ld) {
  if (rb.n <= 0) return;
  e->prof_begin(kFamDecode);
  pgen_records_kernel<<<rb.n, kThreads, 0, e->stream>>>(rb, 0);
  GW_CUDA(cudaGetLastError());
  e->launches++;
  if (any_ld) {
    pgen_records_kernel<<<rb.n, kThreads, 0, e->stream>>>(rb, 1);
    GW_CUDA(cudaGetLastError());
    e->launches++;
  }
  e->prof_end(kFamDecode);
}

void launch_rows_to_doubles(gwsem_engine* e, const uint8_t* d_rows, int64_t pitch, const uint16_t* d_dosage, int n_pad,
                            int n_snps, int n_samples, const int32_t* d_dest_of, const unsigned long long* d_rowmask,
                            int64_t out_stride, double* d_out, double* d_maf) {
  if (n_snps <= 0) return;
  e->d_maf_acc.ensure(2 * (size_t)n_snps);
  GW_CUDA(cudaMemsetAsync(e->d_maf_acc.p, 0, 2 * (size_t)n_snps * sizeof(unsigned long long), e->stream));
  for (int r0 = 0; r0 < n_snps; r0 += 65535) {
    const int rows = std::min(65535, n_snps - r0);
    dim3 grid((n_samples + 255) / 256, rows);
    rows_to_doubles_kernel<<<grid, 256, 0, e->stream>>>(d_rows + (int64_t)r0 * pitch, pitch,
                                                        d_dosage ? d_dosage + (int64_t)r0 * n_pad : nullptr, n_pad, n_samples,
                                                        d_dest_of, d_rowmask, out_stride, d_out + (int64_t)r0 * out_stride,
                                                        e->d_maf_acc.p + 2 * (size_t)r0);
    GW_CUDA(cudaGetLastError());
    e->launches++;
  }
  maf_finish_units_kernel<<<(n_snps + 255) / 256, 256, 0, e->stream>>>(e->d_maf_acc.p, n_snps, d_maf);
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

void launch_bgen_blocks(gwsem_engine* e, const BgenBatch& bb) {
  if (bb.n <= 0) return;
  e->prof_begin(kFamDecode);
  bgen_blocks_kernel<<<bb.n, kThreads, 0, e->stream>>>(bb);
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamDecode);
  e->launches++;
}

}  // namespace gwsem
