// Genotype-class sums on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// The same exact integer contraction as stats_imma.cu -- one-hot genotype classes [SNPs x people] times
// the features as four balanced base-256 digits [people x 4F], int8 x int8 -> int32 -- but issued as
// tcgen05.mma kind::i8 with the one-hot A operand written straight to tensor memory:
//
//   thread t of the CTA = SNP snp0 + t = TMEM lane t.  Per pair of 32-person blocks it reads the 16 packed
//   bytes of its SNP from the staged tile, expands them with PRMT (the packed word is the selector, see
//   stats_imma.cu) into the het and the hom-ALT indicator rows -- 2 x 16 registers -- and stores them with
//   tcgen05.st into its lane of the two A buffers (M = 128 SNPs, K = 32 people per MMA).  One thread then
//   issues, per block, one MMA per class against the feature digits in shared memory (K-major canonical
//   layout without swizzle, packed on the host) into the two accumulators D_het / D_hom [128 x N] in TMEM.
//   The A buffers are double buffered; a tcgen05.commit -> mbarrier per buffer tells the CTA when the MMAs
//   that read it are done.  The legacy mma.sync path spends 8.4 clk/SMSP per IMMA and overlaps only partly
//   with the integer ALU work of the expansion (scripts/mma_rate.cu); here the tensor work is asynchronous
//   and the expansion is the only thing the warps issue.
//
// TMEM columns (32-bit): [0, N) D_het, [N, 2N) D_hom, then A[buffer][class] of 16 columns each (2 blocks x 8).
#include "common.h"
#include "kernels.h"

#include <cstdlib>
#include <string>

namespace gwsem {

namespace {

constexpr int kTcThreads = 128;   // = SNPs per CTA = TMEM lanes
constexpr int kTcGPad = 16;

__device__ __forceinline__ unsigned prmt_raw(unsigned a, unsigned b, unsigned sel) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(addr), "r"(parity)
      : "memory");
}

// 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle: 8 x 16-byte core matrices; LBO = byte distance of the
// two 16-byte K chunks of a block, SBO = byte distance of consecutive 8-column groups (version 1 = sm_100)
__device__ __forceinline__ uint64_t smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_i8_ts(unsigned d_tmem, unsigned a_tmem, uint64_t b_desc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// N = digit columns of the chunk (16 or 64), TILE = people per cp.async stage
template <int N, int TILE, bool PLINK1>
__global__ void __launch_bounds__(kTcThreads)
    class_sums_tc5_kernel(const uint8_t* __restrict__ rows, int64_t pitch, int n_snps, int n_blocks,
                          const int8_t* __restrict__ Bq, const double* __restrict__ inv_scale, int f0, int n_feat,
                          double* __restrict__ sums, unsigned tab1, unsigned tab2, int32_t* __restrict__ counts,
                          int count_slot, int32_t* __restrict__ miss_list, int32_t* __restrict__ miss_n) {
  constexpr int kSnps = kTcThreads;
  constexpr int kGRow = TILE / 4 + kTcGPad;
  constexpr int kBlocks = TILE / 32;
  constexpr int kBBlock = N * 32;                    // feature bytes of one 32-person block
  constexpr int kBStage = kBlocks * kBBlock;
  constexpr int kGStage = kSnps * kGRow;
  constexpr int kStage = kGStage + kBStage;
  constexpr int kCols = 2 * N + 64 <= 128 ? 128 : 256;   // TMEM allocation (power of two)
  constexpr unsigned kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ unsigned tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int snp0 = blockIdx.x * kSnps;
  const int n_tiles = (n_blocks + kBlocks - 1) / kBlocks;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(&tmem_base_sh)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_base_sh;
  const unsigned lane_base = tmem_base + ((unsigned)(warp * 32) << 16);   // this warp's 32 lanes
  const unsigned d_col[2] = {0u, (unsigned)N};
  auto a_col = [&](int buf, int cls) { return (unsigned)(2 * N + (buf * 2 + cls) * 16); };

  // ---- staging (cp.async, one pointer bump per copy; see stats_imma.cu)
  constexpr int kPieces = TILE / 64;
  constexpr int kRowsPerPass = kTcThreads / kPieces;
  constexpr int kPasses = kSnps / kRowsPerPass;
  constexpr int kBCopies = kBStage / 16 / kTcThreads;
  static_assert(kBStage % (16 * kTcThreads) == 0, "staging layout");
  const int gc = tid % kPieces, gr = tid / kPieces;
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned gdst = smem0 + gr * kGRow + gc * 16;
  const unsigned bdst = smem0 + kGStage + tid * 16;
  const int8_t* bsrc = Bq + tid * 16;
  const int last_piece = (int)pitch - 16;
  auto load_stage = [&](int tile, int s) {
    const int off = min(tile * (TILE / 4) + gc * 16, last_piece);
    const unsigned d = gdst + s * kStage;
#pragma unroll
    for (int k = 0; k < kPasses; ++k) {
      const uint8_t* p = rows + (int64_t)min(snp0 + gr + k * kRowsPerPass, n_snps - 1) * pitch + off;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + k * kRowsPerPass * kGRow), "l"(p) : "memory");
    }
    const int8_t* q = bsrc + (int64_t)tile * kBStage;
    const unsigned db = bdst + s * kStage;
#pragma unroll
    for (int k = 0; k < kBCopies; ++k)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db + k * kTcThreads * 16), "l"(q + k * kTcThreads * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  unsigned mdet = 0u;      // missing-call detector (see stats_imma.cu)
  int a_stage = 0;         // pairs of blocks expanded so far; buffer = a_stage & 1
  load_stage(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int s = tile & 1;
    if (tile + 1 < n_tiles) {
      load_stage(tile + 1, s ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // the MMAs read the feature digits through the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint8_t* Gt = smem + s * kStage + tid * kGRow;
    const unsigned b_tile = smem0 + s * kStage + kGStage;
#pragma unroll 1
    for (int kp = 0; kp < kBlocks / 2; ++kp, ++a_stage) {
      const int buf = a_stage & 1;
      const uint4 w4 = *reinterpret_cast<const uint4*>(Gt + kp * 16);
      const unsigned w[4] = {w4.x, w4.y, w4.z, w4.w};
      unsigned ah[16], ao[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned x = w[q], x2 = x >> 2, xh = x >> 16, xh2 = x >> 18;
        mdet |= PLINK1 ? (~x & (x * 2u)) : (x & (x * 2u));
        ah[4 * q + 0] = prmt_raw(tab1, tab1, x);
        ah[4 * q + 1] = prmt_raw(tab1, tab1, x2);
        ah[4 * q + 2] = prmt_raw(tab1, tab1, xh);
        ah[4 * q + 3] = prmt_raw(tab1, tab1, xh2);
        ao[4 * q + 0] = prmt_raw(tab2, tab2, x);
        ao[4 * q + 1] = prmt_raw(tab2, tab2, x2);
        ao[4 * q + 2] = prmt_raw(tab2, tab2, xh);
        ao[4 * q + 3] = prmt_raw(tab2, tab2, xh2);
      }
      if (a_stage >= 2) {   // the MMAs that read this buffer two pairs ago must be done
        mbar_wait(&mbar[buf], (unsigned)(((a_stage >> 1) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      tmem_st16(lane_base + a_col(buf, 0), ah);
      tmem_st16(lane_base + a_col(buf, 1), ao);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t bd = smem_desc(b_tile + (2 * kp + h) * kBBlock, N * 16, 128);
          const unsigned acc = (a_stage > 0 || h > 0) ? 1u : 0u;
          mma_i8_ts(tmem_base + d_col[0], tmem_base + a_col(buf, 0) + h * 8, bd, kIdesc, acc);
          mma_i8_ts(tmem_base + d_col[1], tmem_base + a_col(buf, 1) + h * 8, bd, kIdesc, acc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&mbar[buf]))
                     : "memory");
      }
    }
    // the next iteration's cp.async overwrites the other stage: its MMAs (previous tile) are older than the
    // two pairs waited for below, so draining the last two commits covers them
    {
      const int last = a_stage - 1;
      mbar_wait(&mbar[last & 1], (unsigned)((last >> 1) & 1));
      if (last >= 1) mbar_wait(&mbar[(last - 1) & 1], (unsigned)(((last - 1) >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __syncthreads();
  }

  // ---- epilogue: this thread's SNP, both classes, N digit columns each (the one-hot entries are -1)
  const int snp = snp0 + tid;
  if (miss_list && (mdet & 0xAAAAAAAAu) != 0u && snp < n_snps) miss_list[atomicAdd(miss_n, 1)] = snp;
#pragma unroll
  for (int cls = 0; cls < 2; ++cls) {
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
      int v[16];
      tmem_ld16(lane_base + d_col[cls] + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (snp < n_snps) {
#pragma unroll
        for (int fi = 0; fi < 4; ++fi) {
          const int slot = c0 / 4 + fi, f = f0 + slot;
          if (counts && slot == count_slot) {
            int32_t* c = counts + 8 * (int64_t)snp;
            c[0 + cls] = -v[4 * fi + 0];   // included people
            c[3 + cls] = -v[4 * fi + 1];   // people kept by rowFilter
          }
          if (f < n_feat) {
            const double lo = (double)(-v[4 * fi + 0]) + 256.0 * (double)(-v[4 * fi + 1]);
            const double hi = (double)(-v[4 * fi + 2]) + 256.0 * (double)(-v[4 * fi + 3]);
            sums[((int64_t)snp * 2 + cls) * n_feat + f] = (lo + 65536.0 * hi) * inv_scale[f];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
  }
}

template <int N, int TILE>
void launch_tc5(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1, int n_blocks,
                const int8_t* Bq, const double* inv_scale, int f0, int n_feat, double* d_sums, int32_t* d_counts,
                int count_slot, int32_t* d_miss_list, int32_t* d_miss_n) {
  const size_t smem = 2 * ((size_t)kTcThreads * (TILE / 4 + kTcGPad) + (size_t)(TILE / 32) * N * 32);
  const int grid = (n_snps + kTcThreads - 1) / kTcThreads;
  if (plink1) {
    auto k = class_sums_tc5_kernel<N, TILE, true>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kTcThreads, smem, e->stream>>>(d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums,
                                             0x00FF0000u, 0x000000FFu, d_counts, count_slot, d_miss_list, d_miss_n);
  } else {
    auto k = class_sums_tc5_kernel<N, TILE, false>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kTcThreads, smem, e->stream>>>(d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums,
                                             0x0000FF00u, 0x00FF0000u, d_counts, count_slot, d_miss_list, d_miss_n);
  }
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

}  // namespace

bool cs_use_tc5() {
  static const bool on = [] {
    const char* v = getenv("GWSEM_CS_KERNEL");
    return !(v && std::string(v) == "imma");
  }();
  return on;
}

// Host side of the tcgen05 digit layout.  Chunks as in imma_pack_features (the buffer has the same size).  Within a
// chunk: [block of 32 people][K chunk of 16 bytes (2)][group of 8 digit columns][column (8)][16 bytes]: the K-major
// canonical layout without swizzle that the shared-memory descriptor names (LBO = N * 16, SBO = 128).  The MMA k
// index of person p of a block follows the expansion: k = 4 c + b with c = 4 (p / 16) + 2 ((p % 16) / 8) + p % 2
// the TMEM column and b = (p % 8) / 2 the byte PRMT produces it in.
static inline size_t tc5_offset(int n_cols, int blk, int col, int p) {
  const int c = 4 * (p / 16) + 2 * ((p % 16) / 8) + (p % 2), k = 4 * c + (p % 8) / 2;
  return (size_t)blk * n_cols * 32 + (size_t)(k / 16) * (n_cols / 8) * 128 + (size_t)(col / 8) * 128 + (size_t)(col % 8) * 16 + k % 16;
}

void tc5_pack_features(const std::vector<std::vector<int64_t>>& q, int n_pad, const std::vector<uint8_t>& inc8,
                       const std::vector<uint8_t>& kept, std::vector<int8_t>& out) {
  const int n_feat = (int)q.size(), cf = imma_chunk_features(n_feat);
  const int ncol = cf * 4;
  out.assign(imma_feature_bytes(n_feat, n_pad), 0);
  const size_t chunk_bytes = out.size() / ((n_feat + cf - 1) / cf);
  for (int f = 0; f < n_feat; ++f) {
    int8_t* base = out.data() + (size_t)(f / cf) * chunk_bytes;
    const int fi = f % cf;
    for (int i = 0; i < n_pad; ++i) {
      int64_t v = q[f][i];
      for (int d = 0; d < 4; ++d) {
        const int64_t digit = ((v + 128) & 255) - 128;
        base[tc5_offset(ncol, i / 32, fi * 4 + d, i % 32)] = (int8_t)digit;
        v = (v - digit) >> 8;
      }
      if (v != 0) fail("feature %d does not fit four balanced base-256 digits", f);
    }
  }
  if (imma_has_counts(n_feat)) {
    int8_t* base = out.data() + (size_t)(n_feat / cf) * chunk_bytes;
    const int fi = n_feat % cf;
    for (int i = 0; i < n_pad; ++i) {
      base[tc5_offset(ncol, i / 32, fi * 4 + 0, i % 32)] = inc8[i] ? 1 : 0;
      base[tc5_offset(ncol, i / 32, fi * 4 + 1, i % 32)] = kept[i] ? 1 : 0;
    }
  }
}

void launch_class_sums_tc5(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                           const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_feat,
                           double* d_sums, int32_t* d_counts, int32_t* d_miss_list, int32_t* d_miss_n) {
  if (n_snps <= 0 || n_feat <= 0) return;
  if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(d_packed) & 15))
    fail("packed genotype rows must be 16-byte aligned with a pitch that is a multiple of 16 (pitch=%lld)",
         (long long)pitch);
  const int cf = imma_chunk_features(n_feat), chunks = (n_feat + cf - 1) / cf;
  const size_t chunk_bytes = imma_feature_bytes(n_feat, pl.n_pad) / chunks;
  const int n_live = (pl.n_pad / 32 + 1) / 2 * 2;
  e->prof_begin(kFamClassSums);
  for (int f0 = 0, c = 0; f0 < n_feat; f0 += cf, ++c) {
    const int8_t* Bq = d_featq + (size_t)c * chunk_bytes;
    const bool last = f0 + cf >= n_feat;
    int32_t* d_cnt = (last && d_counts && imma_has_counts(n_feat)) ? d_counts : nullptr;
    int32_t* d_ml = d_cnt ? d_miss_list : nullptr;
    int32_t* d_mn = d_cnt ? d_miss_n : nullptr;
    const int count_slot = n_feat % cf;
    if (cf == 4) launch_tc5<16, 512>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn);
    else launch_tc5<64, 256>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn);
  }
  e->prof_end(kFamClassSums);
}

}  // namespace gwsem
