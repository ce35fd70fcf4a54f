// Per-SNP sufficient statistics on packed hardcalls.
//
// Every statistic the WLS path needs from the genotype column (SURVEY.md section 8a rows
// O1, O2, O4) is a sum over people of g^a * h, with h a SNP-independent per-person
// feature (a monomial of phenotype / moderator columns).  For hardcalls g takes three
// values, so   sum_i g_i^a h_i = 1^a * S1[h] + 2^a * S2[h],   S_c[h] = sum over people
// with genotype class c.  The hot kernel therefore computes genotype-class sums of a small
// feature matrix, straight from the 2-bit records -- no double column is materialised.
//
//   class_sums_kernel   classes 1 and 2, FP64 accumulate (bound: FP64 pipe; genotype bytes
//                       are read once from HBM through a cp.async.bulk (TMA) + mbarrier ring)
//   count_missing_kernel  class counts by popcount and the (rare) missing class, one warp per SNP
#include <cstdlib>

#include "common.h"
#include "fit_program.h"
#include "kernels.h"

namespace gwsem {

constexpr int kRowPad = 16;  // bytes added to each staged genotype row: rows land in different banks

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// PLINK 1 .bed 2-bit codes -> PLINK 2 codes, 16 people at a time
// (00->10, 01->11, 10->01, 11->00; reference src/pgenlib_read.cc:1979-1992).
__device__ __forceinline__ uint32_t bed_to_plink2(uint32_t w) {
  const uint32_t hi = w & 0xAAAAAAAAu, lo = w & 0x55555555u;
  return (~hi & 0xAAAAAAAAu) | ((hi >> 1) ^ lo);
}
__device__ __forceinline__ unsigned long long bed_to_plink2(unsigned long long w) {
  const unsigned long long hi = w & 0xAAAAAAAAAAAAAAAAull, lo = w & 0x5555555555555555ull;
  return (~hi & 0xAAAAAAAAAAAAAAAAull) | ((hi >> 1) ^ lo);
}


// Grid: one CTA per WARPS*S SNP rows.  Warp w accumulates, for its S rows and F features,
//   a1[s][f] = sum_{i: code_i == 1} feat[f][i],   a2[s][f] = sum_{i: code_i == 2} feat[f][i]
// lane l taking people l, l+32, ...  The people axis is cut into TILE-person tiles; a ring of
// STAGES shared-memory stages is filled by cp.async.bulk (TMA) copies -- the 2-bit rows of the
// CTA's SNPs, the F feature rows and the include bytes of the tile -- and signalled through
// mbarriers.  All 32 lanes read the same 32-person word of a genotype row (shared-memory
// broadcast) and pick their own 2 bits; the adds are DFMAs with 0/1 weights.
template <int S, int F, int WARPS, int TILE, int STAGES, bool PLINK1>
__global__ void __launch_bounds__(WARPS * 32 + 32, (S * F <= 24 && WARPS <= 4) ? 3 : 1)
    class_sums_kernel(const uint8_t* __restrict__ packed, int64_t pitch, int n_snps, int n_src,
                      const uint8_t* __restrict__ inc8, const double* __restrict__ feat, int n_pad,
                      int feat_stride_out, double* __restrict__ sums) {
  constexpr int kRows = WARPS * S;
  constexpr int kTileBytes = TILE / 4;
  constexpr int kRowStride = kTileBytes + kRowPad;
  constexpr int kGenoBytes = kRows * kRowStride;
  constexpr int kFeatBytes = F * TILE * 8;
  constexpr int kStageBytes = kGenoBytes + kFeatBytes + TILE;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // STAGES "tile landed" barriers
  uint64_t* empty = full + STAGES;                     // STAGES "tile consumed" barriers
  uint8_t* tiles = smem + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRows;
  const int rows_here = min(kRows, n_snps - row0);
  const int n_tiles = (n_src + TILE - 1) / TILE;
  const int64_t row_bytes = (int64_t)((n_src + 3) / 4 + 15) / 16 * 16;  // bytes worth copying per row (<= pitch)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int stage = t % STAGES;
    const int64_t off = (int64_t)t * kTileBytes;
    const uint32_t gbytes = (uint32_t)min((int64_t)kTileBytes, row_bytes - off);
    const int people = min(TILE, n_pad - t * TILE);  // n_pad is a multiple of 32
    mbar_expect_tx(&full[stage], gbytes * rows_here + (uint32_t)people * (8 * F + 1));
    uint8_t* dst = tiles + (size_t)stage * kStageBytes;
    for (int r = 0; r < rows_here; ++r)
      bulk_g2s(dst + r * kRowStride, packed + (int64_t)(row0 + r) * pitch + off, gbytes, &full[stage]);
    for (int f = 0; f < F; ++f)
      bulk_g2s(dst + kGenoBytes + f * TILE * 8, feat + (int64_t)f * n_pad + (int64_t)t * TILE, people * 8, &full[stage]);
    bulk_g2s(dst + kGenoBytes + kFeatBytes, inc8 + (int64_t)t * TILE, people, &full[stage]);
  };
  if (warp == WARPS) {
    // producer warp: one lane keeps the ring full; consumers never meet at a CTA-wide barrier
    if (lane == 0)
      for (int t = 0; t < n_tiles; ++t) {
        if (t >= STAGES) mbar_wait(&empty[t % STAGES], ((t / STAGES) - 1) & 1);
        issue(t);
      }
    return;
  }

  double a1[S][F], a2[S][F];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int f = 0; f < F; ++f) a1[s][f] = a2[s][f] = 0.0;

  const int sh = 2 * (lane & 15);
  const int half = lane >> 4;
  for (int t = 0; t < n_tiles; ++t) {
    const int stage = t % STAGES;
    mbar_wait(&full[stage], (t / STAGES) & 1);
    const uint8_t* stage_base = tiles + (size_t)stage * kStageBytes;
    const uint8_t* base = stage_base + warp * S * kRowStride + half * 4;
    const double* fs = reinterpret_cast<const double*>(stage_base + kGenoBytes) + lane;
    const uint8_t* is = stage_base + kGenoBytes + kFeatBytes + lane;
    const int people = min(TILE, n_src - t * TILE);
    const int blocks = (people + 31) >> 5;
#pragma unroll 1
    for (int b = 0; b < blocks; ++b) {
      double h[F];
#pragma unroll
      for (int f = 0; f < F; ++f) h[f] = fs[f * TILE + b * 32];
      const uint32_t inc = is[b * 32];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(base + s * kRowStride + b * 8);
        if (PLINK1) w = bed_to_plink2(w);
        const uint32_t code = (w >> sh) & inc;
        // class indicators as doubles (1.0 / 0.0: only the high word differs) feeding DFMAs;
        // nvcc turns both "if (p) a += h" and predicated PTX adds into branches or FSEL pairs
        const double w1 = __hiloint2double(code == 1u ? 0x3FF00000 : 0, 0);
        const double w2 = __hiloint2double(code == 2u ? 0x3FF00000 : 0, 0);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          a1[s][f] = fma(w1, h[f], a1[s][f]);
          a2[s][f] = fma(w2, h[f], a2[s][f]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // this warp is done with the stage
  }

#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int f = 0; f < F; ++f) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a1[s][f] += __shfl_xor_sync(0xffffffffu, a1[s][f], o);
        a2[s][f] += __shfl_xor_sync(0xffffffffu, a2[s][f], o);
      }
    }
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int row = row0 + warp * S + s;
      if (row < n_snps) {
        double* o = sums + (int64_t)row * 2 * feat_stride_out;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          o[f] = a1[s][f];
          o[feat_stride_out + f] = a2[s][f];
        }
      }
    }
  }
}

template <int S, int F, int WARPS, int TILE, int STAGES>
static void launch_chunk(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                         const PeopleLayout& pl, const double* d_feat, int n_feat_total, double* d_sums) {
  constexpr int kRows = WARPS * S;
  const size_t smem = 128 + (size_t)STAGES * (kRows * (TILE / 4 + kRowPad) + F * TILE * 8 + TILE);
  const int grid = (n_snps + kRows - 1) / kRows;
  if (plink1) {
    auto k = class_sums_kernel<S, F, WARPS, TILE, STAGES, true>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, WARPS * 32 + 32, smem, e->stream>>>(d_packed, pitch, n_snps, pl.n_src, pl.d_inc8, d_feat, pl.n_pad,
                                                   n_feat_total, d_sums);
  } else {
    auto k = class_sums_kernel<S, F, WARPS, TILE, STAGES, false>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, WARPS * 32 + 32, smem, e->stream>>>(d_packed, pitch, n_snps, pl.n_src, pl.d_inc8, d_feat, pl.n_pad,
                                                   n_feat_total, d_sums);
  }
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

void launch_class_sums(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                       const PeopleLayout& pl, const double* d_feat, int n_feat, double* d_sums) {
  if (n_snps <= 0 || n_feat <= 0) return;
  if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(d_packed) & 15))
    fail("packed genotype rows must be 16-byte aligned with a pitch that is a multiple of 16 (pitch=%lld)",
         (long long)pitch);
  e->prof_begin(kFamClassSums);
  int f0 = 0;
  while (f0 < n_feat) {  // widest chunks first
    const int left = n_feat - f0;
    const double* feat = d_feat + (size_t)f0 * pl.n_pad;
    double* out = d_sums + f0;
    //                          S  F  WARPS TILE STAGES
    if (left >= 8) { launch_chunk<4, 8, 8, 512, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); f0 += 8; }
    else if (left >= 4) { launch_chunk<8, 4, 4, 1024, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); f0 += 4; }
    else if (left == 3) {
      static const int variant = getenv("GWSEM_CS_VARIANT") ? atoi(getenv("GWSEM_CS_VARIANT")) : 0;  // tuning aid
      switch (variant) {
        case 1: launch_chunk<8, 3, 4, 512, 4>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 2: launch_chunk<8, 3, 4, 512, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 3: launch_chunk<8, 3, 8, 512, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 4: launch_chunk<4, 3, 8, 512, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 5: launch_chunk<4, 3, 4, 256, 4>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 6: launch_chunk<8, 3, 4, 256, 4>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        case 7: launch_chunk<8, 3, 4, 1024, 2>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
        default: launch_chunk<8, 3, 4, 1024, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); break;
      }
      f0 += 3;
    }
    else if (left == 2) { launch_chunk<8, 2, 4, 1024, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); f0 += 2; }
    else { launch_chunk<8, 1, 4, 1024, 3>(e, d_packed, pitch, n_snps, plink1, pl, feat, n_feat, out); f0 += 1; }
  }
  e->prof_end(kFamClassSums);
}

// One warp per SNP row: class counts by popcount over included people and the sums of the
// person-major features over the missing class.  Lanes scan 32-person words; a word with
// missing calls is handled by the whole warp (lane f accumulates feature f, f+32, ...), people
// in index order, so the result does not depend on scheduling.
// LIGHT: the het / hom-ALT counts come from the tensor-core class-sum kernel (indicator columns,
// stats_imma.cu); only the missing class is counted here, and words without a missing call cost
// three logic operations.
template <bool PLINK1, bool LIGHT>
__global__ void __launch_bounds__(128)
    count_missing_kernel(const uint8_t* __restrict__ packed, int64_t pitch, int n_snps, int n_pad,
                         const unsigned long long* __restrict__ incmask,
                         const unsigned long long* __restrict__ rowmask, const double* __restrict__ featT,
                         int n_featm, int n_featm_pad, int32_t* __restrict__ counts, double* __restrict__ miss,
                         const int32_t* __restrict__ list, const int32_t* __restrict__ list_n) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_snps) return;
  if (list) {   // only the rows the class-sum kernel found a missing call in (any order: rows are independent)
    if (row >= *list_n) return;
    row = list[row];
  }
  const unsigned long long* words = reinterpret_cast<const unsigned long long*>(packed + (int64_t)row * pitch);
  const int n_words = n_pad >> 5;
  int c1 = 0, c2 = 0, c3 = 0, r1 = 0, r2 = 0, r3 = 0;
  constexpr int kMaxAcc = 8;  // up to 256 missing-class features
  double acc[kMaxAcc];
#pragma unroll
  for (int k = 0; k < kMaxAcc; ++k) acc[k] = 0.0;
  constexpr int kUnroll = 4;  // 4 independent 8-byte loads per lane in flight
  for (int wbase = 0; wbase < n_words; wbase += 32 * kUnroll) {
    unsigned long long wv[kUnroll], iv[kUnroll], rv[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int wi = wbase + 32 * k + lane;
      const bool ok = wi < n_words;
      wv[k] = ok ? __ldg(&words[wi]) : 0ull;
      iv[k] = ok ? __ldg(&incmask[wi]) : 0ull;
      rv[k] = ok ? __ldg(&rowmask[wi]) : 0ull;
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int w0 = wbase + 32 * k;
      if (w0 >= n_words) break;
      unsigned long long w = wv[k];
      if (PLINK1) w = bed_to_plink2(w);
      const unsigned long long lo = w & 0x5555555555555555ull, hi = (w >> 1) & 0x5555555555555555ull;
      const unsigned long long m3 = lo & hi & iv[k];
      if (LIGHT) {
        const unsigned long long m3r = lo & hi & rv[k];
        if (m3r) { c3 += __popcll(m3); r3 += __popcll(m3r); }
      } else {
        c1 += __popcll(lo & ~hi & iv[k]);
        c2 += __popcll(hi & ~lo & iv[k]);
        c3 += __popcll(m3);
        r1 += __popcll(lo & ~hi & rv[k]);
        r2 += __popcll(hi & ~lo & rv[k]);
        r3 += __popcll(lo & hi & rv[k]);
      }
      unsigned any = __ballot_sync(0xffffffffu, m3 != 0);
      while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        unsigned long long bits = __shfl_sync(0xffffffffu, m3, src);
        const int person0 = (w0 + src) << 5;
        while (bits) {
          const int person = person0 + ((__ffsll((long long)bits) - 1) >> 1);
          bits &= bits - 1;
          const double* f = featT + (int64_t)person * n_featm_pad;
#pragma unroll
          for (int q = 0; q < kMaxAcc; ++q)
            if (lane + 32 * q < n_featm) acc[q] += f[lane + 32 * q];
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    r3 += __shfl_xor_sync(0xffffffffu, r3, o);
  }
  if (lane == 0) {
    if (LIGHT) {
      counts[8 * (int64_t)row + 2] = c3;
      counts[8 * (int64_t)row + 5] = r3;
    } else {
      int4* o = reinterpret_cast<int4*>(counts + 8 * (int64_t)row);
      o[0] = make_int4(c1, c2, c3, r1);
      o[1] = make_int4(r2, r3, 0, 0);
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxAcc; ++k)
    if (lane + 32 * k < n_featm) miss[(int64_t)row * n_featm + lane + 32 * k] = acc[k];
}

void launch_count_missing(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                          const PeopleLayout& pl, const double* d_featT, int n_featm, int n_featm_pad,
                          int32_t* d_counts, double* d_miss, bool light, const int32_t* d_list,
                          const int32_t* d_list_n) {
  if (n_snps <= 0) return;
  if (n_featm > 256) fail("too many missing-class features (%d > 256)", n_featm);
  if (pitch % 8 != 0 || pitch * 4 < pl.n_pad) fail("pitch %lld too small for %d padded people", (long long)pitch, pl.n_pad);
  const int warps = 4;
  const int grid = (n_snps + warps - 1) / warps;
  e->prof_begin(kFamCountMissing);
#define GW_CM(P1, LT)                                                                                              \
  count_missing_kernel<P1, LT><<<grid, warps * 32, 0, e->stream>>>(d_packed, pitch, n_snps, pl.n_pad, pl.d_incmask, \
                                                                   pl.d_rowmask, d_featT, n_featm, n_featm_pad, d_counts, d_miss, d_list, d_list_n)
  if (plink1) { if (light) GW_CM(true, true); else GW_CM(true, false); }
  else { if (light) GW_CM(false, true); else GW_CM(false, false); }
#undef GW_CM
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamCountMissing);
  e->launches++;
}

// R[a][f] from the hardcall class statistics: genotype g is 1 on class 1 and 2 on class 2, so
// sum g^a h = S1[h] + 2^a S2[h]; a = 0 uses totals minus the missing class.  MAF as the reference
// computes it (loader.cpp:386-394): over every person kept by rowFilter, folded to <= 0.5.
__global__ void assemble_moments_kernel(FitProgram fp, int n_kept, int n_snps, const int32_t* __restrict__ counts,
                                        const double* __restrict__ sums, const double* __restrict__ miss,
                                        double* __restrict__ R, int32_t* __restrict__ n_out, double* __restrict__ maf) {
  const int nf = fp.nf;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_snps * 5 * nf) return;
  const int snp = (int)(idx / (5 * nf)), e = (int)(idx % (5 * nf));
  const int a = e / nf, f = e % nf;
  const int32_t* c = counts + 8 * (int64_t)snp;
  const int n = fp.n_incl - c[2];
  double v;
  if (a == 0) {
    v = f == 0 ? (double)n : fp.tot[f] - miss[(int64_t)snp * fp.nfm + (f - 1)];
  } else {
    const double w2 = (double)(1 << a);
    if (f == 0) v = (double)c[0] + w2 * (double)c[1];
    else if (f - 1 < fp.nfd)
      v = sums[(int64_t)snp * 2 * fp.nfd + (f - 1)] + w2 * sums[(int64_t)snp * 2 * fp.nfd + fp.nfd + (f - 1)];
    else v = 0.0;
  }
  R[idx] = v;
  if (e == 0) {
    n_out[snp] = n;
    double m = ((double)c[3] + 2.0 * (double)c[4]) / (2.0 * (double)(n_kept - c[5]));
    if (m > 0.5) m = 1 - m;
    maf[snp] = m;
  }
}

void launch_assemble_moments(gwsem_engine* e, const FitProgram& fp, const PeopleLayout& pl, int n_snps,
                             const int32_t* d_counts, const double* d_sums, const double* d_miss, double* d_R,
                             int32_t* d_n, double* d_maf) {
  if (n_snps <= 0) return;
  const int64_t total = (int64_t)n_snps * 5 * fp.nf;
  assemble_moments_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(fp, pl.n_kept, n_snps, d_counts, d_sums,
                                                                                   d_miss, d_R, d_n, d_maf);
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

// Generic genotype column (dosage / bgen): one CTA per SNP.  Thread t owns features t, t+128, ...
// (up to 2 per thread -> 256 features) and all five powers of g; people are walked in order, their
// genotype value broadcast through shared memory, the person-major feature row read coalesced.
// Deterministic (fixed order), FP64.  This is the fallback path, bound by the feature re-reads.
__global__ void __launch_bounds__(128)
    power_sums_kernel(FitProgram fp, int n_src, int n_pad, const uint8_t* __restrict__ inc8,
                      const double* __restrict__ geno, const double* __restrict__ featT, int n_featm_pad,
                      double* __restrict__ R, int32_t* __restrict__ n_out) {
  __shared__ double g_tile[128];
  __shared__ int cnt_sh;
  const int snp = blockIdx.x, t = threadIdx.x;
  const int nf = fp.nf;
  const double* grow = geno + (int64_t)snp * n_pad;
  double acc[2][5];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int a = 0; a < 5; ++a) acc[k][a] = 0.0;
  int cnt = 0;
  if (t == 0) cnt_sh = 0;
  for (int i0 = 0; i0 < n_src; i0 += 128) {
    const int i = i0 + t;
    double g = __longlong_as_double(0x7FF8000000000000ll);
    if (i < n_src && inc8[i]) g = grow[i];
    __syncthreads();
    g_tile[t] = g;
    __syncthreads();
    const int lim = min(128, n_src - i0);
    for (int j = 0; j < lim; ++j) {
      const double gj = g_tile[j];
      if (!(gj == gj) || isinf(gj)) continue;  // uniform across the CTA
      if (t == 0) ++cnt;
      const double g2 = gj * gj, g3 = g2 * gj, g4 = g2 * g2;
      const double* frow = featT + (int64_t)(i0 + j) * n_featm_pad;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int f = t + 128 * k;
        if (f < nf) {
          const double h = f == 0 ? 1.0 : frow[f - 1];
          acc[k][0] += h;
          acc[k][1] = fma(gj, h, acc[k][1]);
          acc[k][2] = fma(g2, h, acc[k][2]);
          acc[k][3] = fma(g3, h, acc[k][3]);
          acc[k][4] = fma(g4, h, acc[k][4]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int f = t + 128 * k;
    if (f < nf)
#pragma unroll
      for (int a = 0; a < 5; ++a) R[(int64_t)snp * 5 * nf + a * nf + f] = acc[k][a];
  }
  if (t == 0) n_out[snp] = cnt;
}

// The same table with the people axis split over the eight warps of the CTA and S SNPs per CTA
// sharing every feature load: lane = feature (NFG groups of 32), four people in flight per step.
// Each warp owns a contiguous slice of people and sums it in order; the eight partial sums are then
// added in warp order: deterministic.
// NPOW: powers of the genotype kept (5: g^0..g^4 for the fourth-order moments of WLS cumulants; 3: ML reads up to g^2).
template <int S, int NFG, int NPOW>
__global__ void __launch_bounds__(256)
    power_sums_split_kernel(FitProgram fp, int n_snps, int n_src, int n_pad, const uint8_t* __restrict__ inc8,
                            const double* __restrict__ geno, const double* __restrict__ featT, int n_featm_pad,
                            double* __restrict__ R, int32_t* __restrict__ n_out) {
  extern __shared__ double red[];   // [8][NFG * 32][5]
  __shared__ int cnt_sh[8][S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int snp0 = blockIdx.x * S, nf = fp.nf;
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  double acc[S][NFG][NPOW];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int q = 0; q < NFG; ++q)
#pragma unroll
      for (int a = 0; a < NPOW; ++a) acc[s][q][a] = 0.0;
  int cnt[S];
#pragma unroll
  for (int s = 0; s < S; ++s) cnt[s] = 0;
  const int chunk = ((n_src + 7) / 8 + 31) / 32 * 32;
  const int lo = warp * chunk, hi = min(n_src, lo + chunk);
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    double g[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      g[s] = nan;
      if (i < hi && inc8[i] && snp0 + s < n_snps) {
        const double v = geno[(int64_t)(snp0 + s) * n_pad + i];
        if (!isinf(v)) g[s] = v;
      }
    }
    const int lim = min(32, hi - i0);
    for (int j0 = 0; j0 < lim; j0 += 4) {
      double h[4][NFG];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* frow = featT + (int64_t)(i0 + j0 + u) * n_featm_pad;
#pragma unroll
        for (int q = 0; q < NFG; ++q) {
          const int f = q * 32 + lane;
          h[u][q] = (j0 + u < lim && f < nf) ? (f == 0 ? 1.0 : frow[f - 1]) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double gj = __shfl_sync(0xffffffffu, g[s], (j0 + u) & 31);
          const bool ok = gj == gj && j0 + u < lim;
          if (ok) ++cnt[s];
          const double g1 = ok ? gj : 0.0, g2 = g1 * g1, g3 = g2 * g1, g4 = g2 * g2;
#pragma unroll
          for (int q = 0; q < NFG; ++q) {
            const double hv = ok ? h[u][q] : 0.0;
            acc[s][q][0] += hv;
            acc[s][q][1] = fma(g1, hv, acc[s][q][1]);
            acc[s][q][2] = fma(g2, hv, acc[s][q][2]);
            if (NPOW > 3) {
              acc[s][q][3] = fma(g3, hv, acc[s][q][3]);
              acc[s][q][4] = fma(g4, hv, acc[s][q][4]);
            }
          }
        }
      }
    }
  }
  if (lane == 0)
#pragma unroll
    for (int s = 0; s < S; ++s) cnt_sh[warp][s] = cnt[s];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NFG; ++q)
#pragma unroll
      for (int a = 0; a < 5; ++a) red[((warp * NFG + q) * 32 + lane) * 5 + a] = a < NPOW ? acc[s][q][a < NPOW ? a : 0] : 0.0;
    __syncthreads();
    if (snp0 + s < n_snps) {
      for (int e = tid; e < NFG * 32 * 5; e += 256) {
        const int f = e / 5, a = e % 5;
        if (f >= nf) continue;
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[(w * NFG * 32 + f) * 5 + a];
        R[(int64_t)(snp0 + s) * 5 * nf + a * nf + f] = v;
      }
      if (tid == 0) {
        int c = 0;
        for (int w = 0; w < 8; ++w) c += cnt_sh[w][s];
        n_out[snp0 + s] = c;
      }
    }
  }
}

void launch_power_sums(gwsem_engine* e, const FitProgram& fp, const PeopleLayout& pl, int n_snps,
                       const double* d_geno, const double* d_featT, int n_featm_pad, double* d_R, int32_t* d_n, int moment_order) {
  if (n_snps <= 0) return;
  if (fp.nf > 256) fail("too many features for the generic statistics kernel (%d > 256)", fp.nf);
  e->prof_begin(kFamClassSums);
  const int nfg = (fp.nf + 31) / 32;
#define GW_PS(SS, NFG, NPOW)                                                                                         \
  {                                                                                                                  \
    const size_t smem = (size_t)8 * NFG * 32 * 5 * sizeof(double);                                                   \
    GW_CUDA(cudaFuncSetAttribute(power_sums_split_kernel<SS, NFG, NPOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    power_sums_split_kernel<SS, NFG, NPOW><<<(n_snps + SS - 1) / SS, 256, smem, e->stream>>>(                        \
        fp, n_snps, pl.n_src, pl.n_pad, pl.d_inc8, d_geno, d_featT, n_featm_pad, d_R, d_n);                          \
  }
  if (moment_order <= 2) {   // ML: second-order moments only -> g^0 .. g^2, more SNPs per CTA share every feature load
    if (nfg == 1) GW_PS(8, 1, 3)
    else if (nfg == 2) GW_PS(4, 2, 3)
    else if (nfg <= 4) GW_PS(4, 4, 3)
    else GW_PS(2, 8, 3)
  } else {
    if (nfg == 1) GW_PS(4, 1, 5)
    else if (nfg == 2) GW_PS(4, 2, 5)
    else if (nfg <= 4) GW_PS(2, 4, 5)
    else GW_PS(1, 8, 5)
  }
#undef GW_PS
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamClassSums);
  e->launches++;
}

}  // namespace gwsem
