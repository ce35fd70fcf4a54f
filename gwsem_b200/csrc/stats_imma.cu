// Genotype-class sums as an exact integer contraction on the tensor cores.
//
//   sums[snp][class][f] = sum over people i with genotype class(i, snp) == class of h_f(i)
//
// is the product of a one-hot matrix A [2 * SNPs x people] with the feature matrix B [people x F].
// The features are fixed-point integers (gwsem_model::prepare rounds every feature to a grid
// 2^-s_f and the whole engine uses those rounded values), split into four balanced base-256 digits,
// so B is int8, A is 0/1, and mma.sync m16n8k32 (s8 x s8 -> s32) accumulates exactly: no rounding
// anywhere, and the result does not depend on summation order.  The digits' partial sums are
// recombined in FP64 (each below 2^53).
//
// Why mma.sync and not tcgen05 here: tcgen05.mma reads A from shared memory (or TMEM), so the
// one-hot expansion (16x the packed bytes) would have to be written to shared memory first, and
// that store traffic is the bottleneck.  With mma.sync the A fragment is four registers per thread
// that are built in place from two packed bytes with a nibble spread and four PRMT table look-ups
// (PLINK 1 recoding is a different table: free).  The contraction is far from the tensor peak
// either way; the bound is the packed-row stream from HBM and the integer ALU work of the expansion.
//
// Layout.  A warp owns MT m-tiles of 8 SNPs (rows 0-7: het class, rows 8-15: hom-ALT class).  Per
// 32-person block it loads the 8-byte packed word of its SNP (lane>>2), builds the A fragments and
// issues NT MMAs per m-tile against the B fragments (shared by the warp's m-tiles).  Packed rows and
// the int8 feature block are staged in shared memory with cp.async, double buffered.
#include "common.h"
#include "kernels.h"

namespace gwsem {

namespace {

constexpr int kGPad = 16;  // bytes of padding per staged genotype row (keeps 16-byte alignment): 8 rows -> 8 different bank groups

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// prmt.b32 as is (the __byte_perm intrinsic first masks the selector to three bits per nibble; the
// selectors built below never have bit 3 set)
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__device__ __forceinline__ void mma_s8(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NT n-tiles of 8 columns (2 features each), MT m-tiles of 8 SNPs per warp, TILE people per stage.
template <int NT, int MT, int WARPS, int TILE, bool PLINK1>
__global__ void __launch_bounds__(WARPS * 32)
    class_sums_imma_kernel(const uint8_t* __restrict__ rows, int64_t pitch, int n_snps, int n_blocks,
                           const int8_t* __restrict__ Bq, const double* __restrict__ inv_scale, int f0, int n_feat,
                           double* __restrict__ sums, unsigned tab1, unsigned tab2, int32_t* __restrict__ counts,
                           int count_slot, int32_t* __restrict__ miss_list, int32_t* __restrict__ miss_n) {
  constexpr int kSnps = WARPS * MT * 8;
  constexpr int kGRow = TILE / 4 + kGPad;            // staged genotype row, bytes
  constexpr int kBlocks = TILE / 32;                 // 32-person blocks per stage
  constexpr int kBStage = kBlocks * NT * 8 * 32;     // staged feature bytes
  constexpr int kGStage = kSnps * kGRow;
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int kStage = kGStage + kBStage;   // stage s: genotype rows at s * kStage, feature block after them
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int snp0 = blockIdx.x * kSnps;
  const int n_tiles = (n_blocks + kBlocks - 1) / kBlocks;

  // Staging with cp.async: a thread copies the same 16-byte column of kPasses rows of every tile and the
  // same pieces of the (tile-padded, contiguous) feature block, so the addresses are one pointer bump
  // per copy -- the generic index arithmetic cost a third of the kernel's instructions.
  constexpr int kThreads = WARPS * 32;
  constexpr int kPieces = TILE / 64;                  // 16-byte pieces per staged row
  constexpr int kRowsPerPass = kThreads / kPieces;
  constexpr int kPasses = kSnps / kRowsPerPass;
  constexpr int kBCopies = kBStage / 16 / kThreads;
  static_assert(kThreads % kPieces == 0 && kSnps % kRowsPerPass == 0 && kBStage % (16 * kThreads) == 0, "staging layout");
  const int gc = tid % kPieces, gr = tid / kPieces;
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned gdst = smem0 + gr * kGRow + gc * 16;
  const unsigned bdst = smem0 + kGStage + tid * 16;
  const int8_t* bsrc = Bq + tid * 16;
  // No predicates: rows past the batch are clamped to its last row (their results are not stored) and pieces
  // past the pitch to the last piece of the row (the feature digits of those people are zero padding).
  const uint8_t* gp[kPasses];
#pragma unroll
  for (int k = 0; k < kPasses; ++k) gp[k] = rows + (int64_t)min(snp0 + gr + k * kRowsPerPass, n_snps - 1) * pitch;
  const int last_piece = (int)pitch - 16;
  auto load_stage = [&](int tile, int s) {
    const int off = min(tile * (TILE / 4) + gc * 16, last_piece);
    const unsigned d = gdst + s * kStage;
#pragma unroll
    for (int k = 0; k < kPasses; ++k)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + k * kRowsPerPass * kGRow), "l"(gp[k] + off) : "memory");
    const int8_t* q = bsrc + (int64_t)tile * kBStage;
    const unsigned db = bdst + s * kStage;
#pragma unroll
    for (int k = 0; k < kBCopies; ++k)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db + k * kThreads * 16), "l"(q + k * kThreads * 16) : "memory");
    cp_async_commit();
  };

  int acc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[m][j][k] = 0;
  // Missing calls (code 3; PLINK 1: code 1) are rare or absent in imputed data: every packed word is OR-ed
  // into a per-SNP detector (one multiply-by-two and one LOP3), and only the SNPs that trip it are listed
  // for count_missing_kernel.  People masks are not applied here: a false positive only costs time.
  unsigned mdet[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) mdet[m] = 0u;
  // PRMT tables (kernel arguments tab1 / tab2): byte index = 2-bit code, 0xFF (-1 as s8) where the code is
  // the class.  het: PLINK 2 code 1, PLINK 1 code 2; hom ALT: PLINK 2 code 2, PLINK 1 code 0.
  // A selector nibble is four raw bits [c' c' c c]: bits 0-1 are the person's code, bit 2 picks the same
  // table again (both PRMT operands hold it) and bit 3 asks for sign replication, which maps 0x00 / 0xFF
  // to themselves -- so the packed word is the selector as it is, without masking or nibble spreading.

  load_stage(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int s = tile & 1;
    if (tile + 1 < n_tiles) {
      load_stage(tile + 1, s ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint8_t* Gw = smem + s * kStage + (warp * MT * 8 + g) * kGRow + t * 4;
    const uint8_t* Bw = smem + s * kStage + kGStage + g * 32 + t * 8;
#pragma unroll 2
    for (int kp = 0; kp < kBlocks / 2; ++kp) {   // a pair of 32-person blocks = one 16-byte piece of a packed row
      uint2 bf[2][NT];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < NT; ++j) bf[h][j] = *reinterpret_cast<const uint2*>(Bw + ((2 * kp + h) * NT * 8 + j * 8) * 32);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        // 16 people of this lane's SNP: codes 0-7 feed the first block of the pair, 8-15 the second;
        // within a block the even codes are k = 4t.., the odd ones k = 16 + 4t..
        const unsigned w = *reinterpret_cast<const unsigned*>(Gw + m * 8 * kGRow + kp * 16);
        const unsigned w2 = w >> 2, wh = w >> 16, wh2 = w >> 18, ws = w * 2u;
        mdet[m] |= PLINK1 ? (~w & ws) : (w & ws);   // odd bits: low and high bit of a code both set (PLINK 1: low only)
        {
          const unsigned a0 = prmt(tab1, tab1, w), a1 = prmt(tab2, tab2, w);
          const unsigned a2 = prmt(tab1, tab1, w2), a3 = prmt(tab2, tab2, w2);
#pragma unroll
          for (int j = 0; j < NT; ++j) mma_s8(acc[m][j], a0, a1, a2, a3, bf[0][j].x, bf[0][j].y);
        }
        {
          const unsigned a0 = prmt(tab1, tab1, wh), a1 = prmt(tab2, tab2, wh);
          const unsigned a2 = prmt(tab1, tab1, wh2), a3 = prmt(tab2, tab2, wh2);
#pragma unroll
          for (int j = 0; j < NT; ++j) mma_s8(acc[m][j], a0, a1, a2, a3, bf[1][j].x, bf[1][j].y);
        }
      }
    }
    __syncthreads();
  }
  if (miss_list) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      unsigned f = mdet[m] & 0xAAAAAAAAu;
      f |= __shfl_xor_sync(0xffffffffu, f, 1);
      f |= __shfl_xor_sync(0xffffffffu, f, 2);
      const int snp = snp0 + (warp * MT + m) * 8 + g;
      if (t == 0 && f != 0u && snp < n_snps) miss_list[atomicAdd(miss_n, 1)] = snp;
    }
  }
  // the one-hot entries are -1: the accumulators hold the negated sums
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[m][j][k] = -acc[m][j][k];
  // ---- recombine the digits: a thread holds digits (0,1) or (2,3) of one feature per n-tile
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const int snp = snp0 + (warp * MT + m) * 8 + g;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const double w = (t & 1) ? 65536.0 : 1.0;
      double v1 = ((double)acc[m][j][0] + 256.0 * (double)acc[m][j][1]) * w;   // het
      double v2 = ((double)acc[m][j][2] + 256.0 * (double)acc[m][j][3]) * w;   // hom ALT
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
      const int f = f0 + 2 * j + (t >> 1);
      if (counts && 2 * j + (t >> 1) == count_slot && (t & 1) == 0 && snp < n_snps) {
        // indicator columns: digit columns 0 / 1 of this slot = included / kept people
        int32_t* c = counts + 8 * (int64_t)snp;
        c[0] = acc[m][j][0]; c[3] = acc[m][j][1];   // het: included, kept
        c[1] = acc[m][j][2]; c[4] = acc[m][j][3];   // hom ALT
      }
      if ((t & 1) == 0 && snp < n_snps && f < n_feat) {
        const double sc = inv_scale[f];
        sums[((int64_t)snp * 2 + 0) * n_feat + f] = v1 * sc;
        sums[((int64_t)snp * 2 + 1) * n_feat + f] = v2 * sc;
      }
    }
  }
}

template <int NT, int MT, int WARPS, int TILE>
void launch_imma(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1, int n_blocks,
                 const int8_t* Bq, const double* inv_scale, int f0, int n_feat, double* d_sums, int32_t* d_counts,
                 int count_slot, int32_t* d_miss_list, int32_t* d_miss_n) {
  constexpr int kSnps = WARPS * MT * 8;
  const size_t smem = 2 * ((size_t)kSnps * (TILE / 4 + kGPad) + (size_t)(TILE / 32) * NT * 8 * 32);
  const int grid = (n_snps + kSnps - 1) / kSnps;
  if (plink1) {
    auto k = class_sums_imma_kernel<NT, MT, WARPS, TILE, true>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, WARPS * 32, smem, e->stream>>>(d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums, 0x00FF0000u, 0x000000FFu, d_counts, count_slot, d_miss_list, d_miss_n);
  } else {
    auto k = class_sums_imma_kernel<NT, MT, WARPS, TILE, false>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, WARPS * 32, smem, e->stream>>>(d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums, 0x0000FF00u, 0x00FF0000u, d_counts, count_slot, d_miss_list, d_miss_n);
  }
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

}  // namespace

int imma_chunk_features(int n_feat) { return n_feat <= 4 ? 4 : 16; }
// 32-person blocks of the digit layout, padded with zero blocks to whole staging tiles of any kernel
// variant (TILE <= 1024 people): the feature copies need no bounds check
static inline int imma_blocks(int n_pad) { return (n_pad / 32 + 31) / 32 * 32; }

// Host side of the digit layout: chunk c covers features [c * cf, (c + 1) * cf), cf = imma_chunk_features;
// within a chunk: [block of 32 people][column = feature_in_chunk * 4 + digit][position].  Person p of a pair of
// blocks (64 people = 16 packed bytes) sits in 32-bit word t = p / 16 of the piece, which lane t expands: codes
// 0-7 of the word go to the first block, 8-15 to the second, even codes to MMA k = 4t + j, odd ones to
// k = 16 + 4t + j (j = code index / 2); position of k in the block = 8 * ((k % 16) / 4) + 4 * (k / 16) + k % 4
// (the two halves of a thread's B fragment are adjacent).
static inline void imma_position(int i, int& blk, int& pos) {
  const int p = i % 64, t = p / 16, r = p % 16, r8 = r % 8;
  const int k = 16 * (r8 & 1) + 4 * t + r8 / 2;
  blk = (i / 64) * 2 + r / 8;
  pos = 8 * ((k % 16) / 4) + 4 * (k / 16) + k % 4;
}
size_t imma_feature_bytes(int n_feat, int n_pad) {
  const int cf = imma_chunk_features(n_feat), chunks = (n_feat + cf - 1) / cf;
  return (size_t)chunks * imma_blocks(n_pad) * (cf * 4) * 32;
}

bool imma_has_counts(int n_feat) { return n_feat > 0 && n_feat % imma_chunk_features(n_feat) != 0; }

void imma_pack_features(const std::vector<std::vector<int64_t>>& q, int n_pad, const std::vector<uint8_t>& inc8,
                        const std::vector<uint8_t>& kept, std::vector<int8_t>& out) {
  const int n_feat = (int)q.size(), cf = imma_chunk_features(n_feat), chunks = (n_feat + cf - 1) / cf;
  const int n_blocks = imma_blocks(n_pad), ncol = cf * 4;
  out.assign(imma_feature_bytes(n_feat, n_pad), 0);
  for (int f = 0; f < n_feat; ++f) {
    const int c = f / cf, fi = f % cf;
    int8_t* base = out.data() + (size_t)c * n_blocks * ncol * 32;
    for (int i = 0; i < n_pad; ++i) {
      int64_t v = q[f][i];
      int blk, pos;
      imma_position(i, blk, pos);
      for (int d = 0; d < 4; ++d) {
        const int64_t digit = ((v + 128) & 255) - 128;  // balanced digit in [-128, 127]
        base[((size_t)blk * ncol + fi * 4 + d) * 32 + pos] = (int8_t)digit;
        v = (v - digit) >> 8;
      }
      if (v != 0) fail("feature %d does not fit four balanced base-256 digits", f);
    }
  }
  if (imma_has_counts(n_feat)) {  // indicators in the free slot after the last feature
    const int c = n_feat / cf, fi = n_feat % cf;
    int8_t* base = out.data() + (size_t)c * n_blocks * ncol * 32;
    for (int i = 0; i < n_pad; ++i) {
      int blk, pos;
      imma_position(i, blk, pos);
      base[((size_t)blk * ncol + fi * 4 + 0) * 32 + pos] = inc8[i] ? 1 : 0;
      base[((size_t)blk * ncol + fi * 4 + 1) * 32 + pos] = kept[i] ? 1 : 0;
    }
  }
}

void launch_class_sums_imma(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                            const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_feat,
                            double* d_sums, int32_t* d_counts, int32_t* d_miss_list, int32_t* d_miss_n) {
  if (n_snps <= 0 || n_feat <= 0) return;
  if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(d_packed) & 15))
    fail("packed genotype rows must be 16-byte aligned with a pitch that is a multiple of 16 (pitch=%lld)",
         (long long)pitch);
  const int n_blocks = imma_blocks(pl.n_pad), cf = imma_chunk_features(n_feat);
  const int n_live = (pl.n_pad / 32 + 1) / 2 * 2;   // blocks that hold people (whole pairs); the rest is padding
  e->prof_begin(kFamClassSums);
  for (int f0 = 0, c = 0; f0 < n_feat; f0 += cf, ++c) {
    const int8_t* Bq = d_featq + (size_t)c * n_blocks * (cf * 4) * 32;
    // the chunk that holds the indicator columns also writes the class counts
    const bool last = f0 + cf >= n_feat;
    int32_t* d_cnt = (last && d_counts && imma_has_counts(n_feat)) ? d_counts : nullptr;
    const int count_slot = n_feat % cf;
    int32_t* d_ml = d_cnt ? d_miss_list : nullptr;   // the chunk that writes the counts also lists the SNPs with missing calls
    int32_t* d_mn = d_cnt ? d_miss_n : nullptr;
    if (cf == 4) {
      static const int variant = getenv("GWSEM_IMMA_VARIANT") ? atoi(getenv("GWSEM_IMMA_VARIANT")) : 0;  // tuning aid
      switch (variant) {  //          NT MT WARPS TILE   (0.70 / 0.79 / 0.75 ms per 65,536 SNPs x 100k people)
        case 1: launch_imma<2, 4, 4, 512>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn); break;
        case 2: launch_imma<2, 2, 8, 1024>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn); break;
        default: launch_imma<2, 2, 4, 512>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn); break;
      }
    }
    else launch_imma<8, 2, 4, 512>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn);
  }
  e->prof_end(kFamClassSums);
}

}  // namespace gwsem
