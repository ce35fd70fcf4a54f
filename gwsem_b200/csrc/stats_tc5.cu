// Genotype-class sums on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// The same exact integer contraction as stats_imma.cu -- one-hot genotype classes [SNPs x people] times
// the features as four balanced base-256 digits [people x 4F], int8 x int8 -> int32 -- but issued as
// tcgen05.mma kind::i8 with the one-hot A operand written straight to tensor memory:
//
//   SNP snp0 + t = TMEM lane t (M = 128 SNPs per CTA).  Expanding thread t reads, per pair of 32-person blocks, the 16
//   packed bytes of its SNP from the staged tile, expands them with PRMT (the packed word is the selector, see
//   stats_imma.cu) into the het and the hom-ALT indicator rows -- 2 x 16 registers -- and stores them with tcgen05.st
//   into its lane of an A buffer of the ring (K = 32 people per MMA).  With a ring of four buffers two groups of four
//   expanding warps take alternate stages.  Two more warps issue the MMAs, one per genotype class: per block one
//   tcgen05.mma against the feature digits in shared memory (K-major canonical layout without swizzle, packed on the
//   host) into the class's accumulator D_het / D_hom [128 x N] in TMEM, then a tcgen05.commit -> mbarrier that tells
//   the expanding warps when the buffer is free again.  The tile's feature digits arrive by one cp.async.bulk (TMA)
//   per tile, the packed rows by cp.async.  N = 16 / 64 digit columns for the cumulants and ML class sums, 128 / 144
//   for the series path of the WLS marginals statistics (marg_series.cu).  What bounds it: DESIGN.md section 4.
//
// TMEM columns (32-bit): [0, N) D_het, [N, 2N) D_hom, then A[buffer][class] of 16 columns each (2 blocks x 8).
#include "common.h"
#include "kernels.h"

#include <cstdlib>
#include <string>
#include <type_traits>

namespace gwsem {

namespace {

constexpr int kTcThreads = 128;   // SNPs per CTA = TMEM lanes = threads of one expanding group
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

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (UBLKCP in SASS); bytes a multiple of 16, both addresses 16-byte
// aligned; completion is counted on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// the same copy delivered to the same shared-memory offset of every CTA in `mask` of this cluster, one read of L2;
// each destination's own mbarrier (same offset) counts the bytes
__device__ __forceinline__ void tma_bulk_g2s_multicast(unsigned dst, const void* src, unsigned bytes, uint64_t* bar, unsigned short mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "h"(mask)
      : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, unsigned rank) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// Expanding groups: with a ring of four A buffers two groups of four warps take alternate A stages (group g: stages
// g, g + 2, ...; buffers g and g + 2).  One stage is a serial chain for a warp -- expand, wait for the buffer, store,
// tcgen05.wait::st, fence, arrive: about 600 cycles whatever the MMA size -- so one group alone left the tensor pipe
// 46 % busy at 128 columns; two chains in flight feed it twice as fast.
#ifndef GW_TC5_DEBUG
#define GW_TC5_DEBUG 0
#endif
#ifndef GW_TC5_GROUPS
#define GW_TC5_GROUPS 2
#endif
template <int N>
struct TcShape {
  static constexpr int kCols = 2 * N + 3 * 32 <= 128 ? 128 : (2 * N + 3 * 32 <= 256 ? 256 : 512);
  static constexpr int kBufs = (kCols - 2 * N) / 32 < 4 ? (kCols - 2 * N) / 32 : 4;
  static constexpr int kGroups = kBufs == 4 ? GW_TC5_GROUPS : 1;
  static constexpr int kThreads = kTcThreads * kGroups + 64;   // expanding groups + two MMA warps
};

template <int N, int TILE, bool PLINK1>
__global__ void __launch_bounds__(TcShape<N>::kThreads)
    class_sums_tc5_kernel(const uint8_t* __restrict__ rows, int64_t pitch, int n_snps, int n_blocks,
                          const int8_t* __restrict__ Bq, const double* __restrict__ inv_scale, int f0, int n_feat,
                          double* __restrict__ sums, unsigned tab1, unsigned tab2, int32_t* __restrict__ counts,
                          int count_slot, int32_t* __restrict__ miss_list, int32_t* __restrict__ miss_n,
                          unsigned m30, unsigned m16, unsigned m14, int dbg) {
#if !GW_TC5_DEBUG
  dbg = 0;   // the GWSEM_TC5_DBG switches (development aid) are compiled out unless built with -DGW_TC5_DEBUG=1
#endif
  constexpr int kSnps = kTcThreads;
  constexpr int kGRow = TILE / 4 + kTcGPad;
  constexpr int kBlocks = TILE / 32;
  constexpr int kBBlock = N * 32;                    // feature bytes of one 32-person block
  constexpr int kBStage = kBlocks * kBBlock;
  constexpr int kGStage = kSnps * kGRow;
  constexpr int kStage = kGStage + kBStage;
  constexpr int kCols = TcShape<N>::kCols;     // TMEM allocation (power of two)
  constexpr int kBufs = TcShape<N>::kBufs;     // A buffers in tensor memory: 3 (N = 16) or 4 (a ring of 6 at N = 128 measured slower: 18.5 vs 16.0 ms per 33 launches)
  constexpr int kGroups = TcShape<N>::kGroups; // expanding groups of four warps
  constexpr int kExp = kTcThreads * kGroups;   // expanding threads
  static_assert(kBufs == 3 || kBufs == 4, "A ring");
  static_assert(2 * N + kBufs * 32 <= 512, "tensor memory");
  constexpr unsigned kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kBufs], empty_bar[kBufs], tile_bar[2], free_bar[2];
  __shared__ unsigned tmem_base_sh;
  __shared__ unsigned mdet_sh[TcShape<N>::kGroups][kTcThreads];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int snp0 = blockIdx.x * kSnps;
  // Thread-block cluster (1 when launched without one): the CTAs of a cluster walk the same feature digits, so every
  // tile is read from L2 once and multicast into all of them (each CTA fetches 1 / csize of it for everybody).  L2 -> SM
  // traffic of the digits was what bounded the kernel (5.2 TB/s, ncu r02b).
  unsigned crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int n_tiles = (n_blocks + kBlocks - 1) / kBlocks;
  constexpr int kPairs = kBlocks / 2;                    // A stages (pairs of blocks) per tile

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(&tmem_base_sh)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int b = 0; b < kBufs; ++b) {
      mbar_init(&full_bar[b], kTcThreads / 32);   // one arrival per expanding warp of the stage's group, once its lanes of the A buffer are stored
      mbar_init(&empty_bar[b], 2);           // tcgen05.commit of each of the two MMA warps: the MMAs that read the buffer are done
    }
    mbar_init(&tile_bar[0], 1);              // staging of the feature digits: thread 0 announces the bytes of its bulk copy
    mbar_init(&tile_bar[1], 1);
    mbar_init(&free_bar[0], csize);          // every CTA of the cluster says when this stage of its shared memory may be overwritten
    mbar_init(&free_bar[1], csize);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (csize > 1) cluster_sync_all();   // every CTA's barriers exist before anybody signals or multicasts into them
  const unsigned tmem_base = tmem_base_sh;
  const unsigned smem0 = (unsigned)__cvta_generic_to_shared(smem);
  constexpr unsigned kDHet = 0u, kDHom = (unsigned)N, kA0 = 2u * N;   // A[buf][cls] at kA0 + (2 buf + cls) * 16

  if (warp >= 4 * kGroups) {
    // ---- two MMA warps (one lane each), one per genotype class: per A stage one MMA per 32-person block into the class's
    // own accumulator, then a commit; the buffer is free when both commits have arrived.  One thread issuing all four
    // MMAs of a stage was the kernel's critical path: about 150 dependent instructions per stage, 500 cycles, whatever
    // the expanding warps and the tensor pipe did (ncu r02b: IPC 0.32, tensor pipe 51 % busy at 144 columns).
    const int cls = warp - 4 * kGroups;
    if (lane == 0) {
      const int total = n_tiles * kPairs;
      const unsigned d_acc = tmem_base + (cls ? kDHom : kDHet);
      int buf = 0;
      unsigned phase = 0;
      for (int a = 0, tile = 0, kp = 0; a < total; ++a) {
        if (kp == 0) mbar_wait(&tile_bar[tile & 1], (unsigned)((tile >> 1) & 1));   // the feature digits of this tile have landed
        mbar_wait(&full_bar[buf], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned b_blk = smem0 + (tile & 1) * kStage + kGStage + (2 * kp) * kBBlock;
        const unsigned a_buf = tmem_base + kA0 + (2 * buf + cls) * 16;
        if (!(dbg & 1)) {   // development aid (GWSEM_TC5_DBG): no MMAs
          mma_i8_ts(d_acc, a_buf, smem_desc(b_blk, N * 16, 128), kIdesc, a > 0 ? 1u : 0u);
          mma_i8_ts(d_acc, a_buf + 8, smem_desc(b_blk + kBBlock, N * 16, 128), kIdesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&empty_bar[buf]))
                     : "memory");
        if (++buf == kBufs) { buf = 0; phase ^= 1u; }
        if (++kp == kPairs) { kp = 0; ++tile; }
      }
    }
    __syncwarp();
  } else {
    // ---- expanding warps: thread t of group g = SNP t = TMEM lane t (a warp reaches the lane quarter warp % 4)
    const int grp = tid / kTcThreads, t = tid % kTcThreads;
    const unsigned lane_base = tmem_base + ((unsigned)((warp & 3) * 32) << 16);
    // Staging.  The tile's feature digits are one contiguous kBStage-byte piece of the chunk: thread 0 hands it to the
    // TMA engine (cp.async.bulk, UBLKCP in SASS; the MMA warp waits for it on tile_bar).  The packed rows stay on
    // cp.async with eight threads per 128-byte row piece: one bulk copy per row and thread (128 small copies per tile)
    // was measured 35 % slower for the whole kernel.
    constexpr int kPieces = TILE / 64;
    constexpr int kRowsPerPass = kExp / kPieces;
    constexpr int kPasses = kSnps / kRowsPerPass;
    static_assert(kBStage % 16 == 0 && kSnps % kRowsPerPass == 0, "staging layout");
    const int gc = tid % kPieces, gr = tid / kPieces;
    const unsigned gdst = smem0 + gr * kGRow + gc * 16;
    const int last_piece = (int)pitch - 16;
    const uint8_t* gp[kPasses];   // rows past the batch are clamped to its last row (not stored), pieces past the pitch to the last piece
#pragma unroll
    for (int k = 0; k < kPasses; ++k) gp[k] = rows + (int64_t)min(snp0 + gr + k * kRowsPerPass, n_snps - 1) * pitch;
    auto load_stage = [&](int tile, int s) {
      const int off = min(tile * (TILE / 4) + gc * 16, last_piece);
      const unsigned d = gdst + s * kStage;
      if (!(dbg & 16)) {   // development aid (GWSEM_TC5_DBG): no packed rows
#pragma unroll
        for (int k = 0; k < kPasses; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + k * kRowsPerPass * kGRow), "l"(gp[k] + off) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (tid == 0) {
        // the MMAs that read this stage's digits are done (the empty-barrier wait before this call covers the previous tile)
        if (dbg & 8) {   // development aid (GWSEM_TC5_DBG): no feature digits (what does their ingest cost?)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&tile_bar[s])) : "memory");
          return;
        }
        mbar_arrive_expect_tx(&tile_bar[s], (unsigned)kBStage);
        if (csize == 1) {
          tma_bulk_g2s(smem0 + s * kStage + kGStage, Bq + (int64_t)tile * kBStage, (unsigned)kBStage, &tile_bar[s]);
        } else {
          // tell every CTA of the cluster that this stage of mine is free, wait until all of theirs are, then fetch my
          // share of the tile for everybody
          for (unsigned r = 0; r < csize; ++r) mbar_arrive_remote(&free_bar[s], r);
          mbar_wait(&free_bar[s], (unsigned)((tile >> 1) & 1));
          const unsigned part = (unsigned)kBStage / csize;
          tma_bulk_g2s_multicast(smem0 + s * kStage + kGStage + crank * part, Bq + (int64_t)tile * kBStage + crank * part, part,
                                 &tile_bar[s], (unsigned short)((1u << csize) - 1u));
        }
      }
    };
    auto expanders_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kExp) : "memory"); };

    unsigned mdet = 0u;      // missing-call detector (see stats_imma.cu)
    // stage of a tile at which the next tile's copies start: the first one of this group at or after kBufs (its
    // empty-barrier wait then covers every MMA of the previous tile)
    constexpr int kIssue0 = kBufs < kPairs ? kBufs : kPairs - 1;
    static_assert(kIssue0 >= kBufs && kIssue0 + kGroups - 1 < kPairs, "tiles shorter than the A ring are not handled");
    static_assert(kPairs % kGroups == 0 && kIssue0 % kGroups == 0, "every group meets every tile at the same stages");
    const int kIssue = kIssue0 + grp;
    const int total = n_tiles * kPairs;   // A stages; stage a uses buffer a % kBufs, phase = parity of a / kBufs
    int a = grp, kp = grp, tile = 0;      // this group's stages: grp, grp + kGroups, ...
    unsigned phase = 0;
    const uint8_t* Gt = nullptr;
    load_stage(0, 0);
    // One A stage with a compile-time buffer: the TMEM addresses and barrier addresses are constants of the unrolled body.
    auto do_stage = [&](auto bufc) {
      constexpr int BUF = decltype(bufc)::value;
      if (kp == grp) {   // this group's first stage of the tile
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        expanders_sync();   // the packed rows landed for everyone; everyone is done with the previous tile's
        Gt = smem + (tile & 1) * kStage + t * kGRow;
        // The copies of the next tile can only start half a tile from now (the other stage is still being read), about
        // 1 us before they are needed: less than a trip to HBM under load.  Ask L2 for those row pieces now.
        if (!(dbg & 4) && tile + 1 < n_tiles) {
          const int off = min((tile + 1) * (TILE / 4) + gc * 16, last_piece);
#pragma unroll
          for (int k = 0; k < kPasses; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp[k] + off));
        }
      }
      const uint4 w4 = *reinterpret_cast<const uint4*>(Gt + kp * 16);
      unsigned ah[16], ao[16];
      // right shifts as multiply-high by a power of two the compiler cannot see (kernel arguments): they go to the
      // FMA pipe, the PRMTs keep the ALU pipe to themselves
#define GW_EXPAND(Q, X)                                                                       \
  {                                                                                           \
    const unsigned x = (X), x2 = __umulhi(x, m30), xh = __umulhi(x, m16), xh2 = __umulhi(x, m14); \
    mdet |= PLINK1 ? (~x & (x * 2u)) : (x & (x * 2u));                                        \
    ah[4 * Q + 0] = prmt_raw(tab1, tab1, x);                                                  \
    ah[4 * Q + 1] = prmt_raw(tab1, tab1, x2);                                                 \
    ah[4 * Q + 2] = prmt_raw(tab1, tab1, xh);                                                 \
    ah[4 * Q + 3] = prmt_raw(tab1, tab1, xh2);                                                \
    ao[4 * Q + 0] = prmt_raw(tab2, tab2, x);                                                  \
    ao[4 * Q + 1] = prmt_raw(tab2, tab2, x2);                                                 \
    ao[4 * Q + 2] = prmt_raw(tab2, tab2, xh);                                                 \
    ao[4 * Q + 3] = prmt_raw(tab2, tab2, xh2);                                                \
  }
      GW_EXPAND(0, w4.x) GW_EXPAND(1, w4.y) GW_EXPAND(2, w4.z) GW_EXPAND(3, w4.w)
#undef GW_EXPAND
      if (a >= kBufs) {   // the MMAs that read this buffer kBufs stages ago must be done (one lane polls for its warp)
        if (lane == 0) mbar_wait(&empty_bar[BUF], phase ^ 1u);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      // That commit also covers every MMA of the previous tile once kp >= kBufs, and all expanding threads left
      // that tile's packed rows at the barrier above: the other shared-memory stage is free.
      if (kp == kIssue && tile + 1 < n_tiles) load_stage(tile + 1, (tile + 1) & 1);
      if (!(dbg & 2)) {   // development aid (GWSEM_TC5_DBG): no stores to tensor memory
        tmem_st16(lane_base + kA0 + (2 * BUF + 0) * 16, ah);
        tmem_st16(lane_base + kA0 + (2 * BUF + 1) * 16, ao);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        unsigned x = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) x ^= ah[q] ^ ao[q];
        mdet |= x & 1u;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      // one arrival per warp (tcgen05.wait::st is warp-wide: every lane's stores are done): 128 per-thread arrivals
      // on one barrier word per stage were a serial cost of their own
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&full_bar[BUF])) : "memory");
      a += kGroups;
      kp += kGroups;
      if (kp >= kPairs) { kp -= kPairs; ++tile; }
    };
    if constexpr (kGroups == 4) {
      static_assert(kBufs == 4, "four groups, one buffer of the ring each");
#pragma unroll 1
      while (a < total) {
        if (grp == 0) do_stage(std::integral_constant<int, 0>{});
        else if (grp == 1) do_stage(std::integral_constant<int, 1>{});
        else if (grp == 2) do_stage(std::integral_constant<int, 2>{});
        else do_stage(std::integral_constant<int, 3>{});
        phase ^= 1u;
      }
    } else if constexpr (kGroups == 2) {
      static_assert(kBufs == 4, "two groups share a ring of four");
#pragma unroll 1
      while (a < total) {
        if (grp == 0) {
          do_stage(std::integral_constant<int, 0>{});
          if (a < total) do_stage(std::integral_constant<int, 2>{});
        } else {
          do_stage(std::integral_constant<int, 1>{});
          if (a < total) do_stage(std::integral_constant<int, 3>{});
        }
        phase ^= 1u;
      }
    } else {
#pragma unroll 1
      while (a < total) {
        do_stage(std::integral_constant<int, 0>{});
        if (a < total) do_stage(std::integral_constant<int, 1>{});
        if (a < total) do_stage(std::integral_constant<int, 2>{});
        if constexpr (kBufs > 3) {
          if (a < total) do_stage(std::integral_constant<int, 3>{});
        }
        phase ^= 1u;
      }
    }
    // all MMAs done: the last kBufs commits
#pragma unroll
    for (int d = 1; d <= kBufs; ++d)
      if (total >= d) mbar_wait(&empty_bar[(total - d) % kBufs], (unsigned)(((total - d) / kBufs) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: this thread's SNP (the one-hot entries are -1); the groups share the classes and column halves
    const int snp = snp0 + t;
    if (kGroups > 1) {
      mdet_sh[grp][t] = mdet;
      expanders_sync();
      if (grp == 0)
        for (int g2 = 1; g2 < kGroups; ++g2) mdet |= mdet_sh[g2][t];
    }
    if (grp == 0 && miss_list && (mdet & 0xAAAAAAAAu) != 0u && snp < n_snps) miss_list[atomicAdd(miss_n, 1)] = snp;
    constexpr int kColParts = kGroups == 4 ? 2 : 1;   // column ranges of an accumulator shared out between groups
    if constexpr (N == kTcNarrowFeatures * 3) {
      // three balanced base-256 digits per feature, digit-major: columns [d * 48, d * 48 + 48) hold digit d of the chunk's
      // 48 features (the SNP-dependent rows of Muthen's A and B need 1e-6, not 2^-30, of the feature's range)
      static_assert(kGroups == 2 || kGroups == 4, "one class per group (pair)");
      const int cls = grp & 1;
#pragma unroll
      for (int g3 = 0; g3 < kTcNarrowFeatures / 16; ++g3) {
        if (kGroups == 4 && (g3 & 1) != (grp >> 1)) continue;   // the two groups of a class share its feature groups
        int v0[16], v1[16], v2[16];
        const unsigned d_base = lane_base + (cls ? kDHom : kDHet) + 16 * g3;
        tmem_ld16(d_base, v0);
        tmem_ld16(d_base + kTcNarrowFeatures, v1);
        tmem_ld16(d_base + 2 * kTcNarrowFeatures, v2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (snp < n_snps) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int f = f0 + 16 * g3 + i;
            if (f < n_feat)
              sums[((int64_t)snp * 2 + cls) * n_feat + f] =
                  ((double)(-v0[i]) + 256.0 * (double)(-v1[i]) + 65536.0 * (double)(-v2[i])) * inv_scale[f];
          }
        }
      }
    } else {
#pragma unroll
    for (int cls = 0; cls < 2; ++cls) {
      if (kGroups >= 2 && cls != (grp & 1)) continue;
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 16) {
        if (kColParts == 2 && (c0 >= N / 2) != (grp >= 2)) continue;
        int v[16];
        tmem_ld16(lane_base + (cls ? kDHom : kDHet) + c0, v);
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
    }   // four-digit chunks
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // nobody leaves while a neighbour may still signal its barriers
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
  }
}

template <int N, int TILE>
void launch_tc5(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1, int n_blocks,
                const int8_t* Bq, const double* inv_scale, int f0, int n_feat, double* d_sums, int32_t* d_counts,
                int count_slot, int32_t* d_miss_list, int32_t* d_miss_n) {
  const size_t smem = 2 * ((size_t)kTcThreads * (TILE / 4 + kTcGPad) + (size_t)(TILE / 32) * N * 32);
  static const int dbg = getenv("GWSEM_TC5_DBG") ? atoi(getenv("GWSEM_TC5_DBG")) : 0;   // development aid: 1 no MMAs, 2 no tensor-memory stores
  // clusters of CTAs share every feature tile through one multicast read of L2 (the series path's launches; the chunk's
  // tile bytes must split into 16-byte granular shares).  GWSEM_TC5_CLUSTER sets the size; default 1 = off: measured, the
  // lock step it forces on the CTAs costs more than the traffic it saves (24 launches: 11.0 ms alone, 14.5 in pairs, 23.0 in fours).
  static const int cl_env = getenv("GWSEM_TC5_CLUSTER") ? atoi(getenv("GWSEM_TC5_CLUSTER")) : 1;
  int cluster = (N >= 128 && (cl_env == 2 || cl_env == 4)) ? cl_env : 1;
  int grid = (n_snps + kTcThreads - 1) / kTcThreads;
  if (grid < cluster) cluster = 1;
  grid = (grid + cluster - 1) / cluster * cluster;   // CTAs past the batch repeat its last rows and store nothing
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TcShape<N>::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster > 1 ? 1 : 0;
  if (plink1) {
    auto k = class_sums_tc5_kernel<N, TILE, true>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GW_CUDA(cudaLaunchKernelEx(&cfg, k, d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums, 0x00FF0000u, 0x000000FFu,
                               d_counts, count_slot, d_miss_list, d_miss_n, 1u << 30, 1u << 16, 1u << 14, dbg));
  } else {
    auto k = class_sums_tc5_kernel<N, TILE, false>;
    GW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GW_CUDA(cudaLaunchKernelEx(&cfg, k, d_packed, pitch, n_snps, n_blocks, Bq, inv_scale, f0, n_feat, d_sums, 0x0000FF00u, 0x00FF0000u,
                               d_counts, count_slot, d_miss_list, d_miss_n, 1u << 30, 1u << 16, 1u << 14, dbg));
  }
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

}  // namespace

// Which contraction kernel serves a model with n_feat class-summed features.  The tcgen05 kernel is bound by the
// tensor-memory store port (the one-hot A operand is 16x the packed bytes; measured 128 B/clk/SM), the mma.sync
// kernel by IMMA issue + ALU: equal at 16 digit columns (0.70 ms per 65,536 SNPs x 100k people), and the tcgen05
// kernel wins 1.6x at 64 columns, where mma.sync needs four times the IMMAs for the same expansion.
// GWSEM_CS_KERNEL=imma|tc5 forces one of them (development aid).
bool cs_use_tc5(int n_feat) {
  const char* v = getenv("GWSEM_CS_KERNEL");   // read when a model is prepared (the digit layout follows the kernel)
  if (v && std::string(v) == "imma") return false;
  if (v && std::string(v) == "tc5") return true;
  return imma_chunk_features(n_feat) > 4;
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

// Chunks of the series path (marg_series.cu packs them in parallel), same K-major layout:
//   wide   kTcWideFeatures features x 4 digits, feature-major columns (4 f + d): 128 columns
//   narrow kTcNarrowFeatures features x 3 digits, digit-major columns (48 d + f): 144 columns
size_t tc5_wide_chunk_bytes(int n_pad) { return (size_t)((n_pad / 32 + 31) / 32 * 32) * (kTcWideFeatures * 4) * 32; }
size_t tc5_narrow_chunk_bytes(int n_pad) { return (size_t)((n_pad / 32 + 31) / 32 * 32) * (kTcNarrowFeatures * 3) * 32; }
void tc5_pack_chunk(const int64_t* q, int fi0, int n_in_chunk, int n_pad, int8_t* chunk_base, bool narrow) {
  // features fi0 .. fi0 + n_in_chunk - 1 of the chunk from q[feature - fi0][n_pad] (distinct features write distinct bytes)
  const int digits = narrow ? 3 : 4, ncol = narrow ? kTcNarrowFeatures * 3 : kTcWideFeatures * 4;
  for (int fi = fi0; fi < fi0 + n_in_chunk; ++fi)
    for (int i = 0; i < n_pad; ++i) {
      int64_t v = q[(size_t)(fi - fi0) * n_pad + i];
      if (v == 0) continue;
      for (int d = 0; d < digits; ++d) {
        const int64_t digit = ((v + 128) & 255) - 128;
        const int col = narrow ? d * kTcNarrowFeatures + fi : fi * 4 + d;
        chunk_base[tc5_offset(ncol, i / 32, col, i % 32)] = (int8_t)digit;
        v = (v - digit) >> 8;
      }
      if (v != 0) fail("series feature does not fit %d balanced base-256 digits", digits);
    }
}

// The series path's class sums: n_wide features (a multiple of 32) in wide chunks, then n_narrow features (a multiple
// of 48) in narrow chunks; d_sums[snp][2][n_wide + n_narrow].  The one-hot expansion of the packed rows and the
// re-read of the rows are shared by 128 / 144 digit columns per launch.
void launch_class_sums_tc5_wide(gwsem_engine* e, const uint8_t* d_packed, int64_t pitch, int n_snps, int plink1,
                                const PeopleLayout& pl, const int8_t* d_featq, const double* d_inv_scale, int n_wide,
                                int n_narrow, double* d_sums) {
  if (n_snps <= 0 || n_wide + n_narrow <= 0) return;
  if (n_wide % kTcWideFeatures != 0 || n_narrow % kTcNarrowFeatures != 0)
    fail("series class sums: %d + %d features (multiples of %d and %d expected)", n_wide, n_narrow, kTcWideFeatures, kTcNarrowFeatures);
  if (pitch % 16 != 0 || (reinterpret_cast<uintptr_t>(d_packed) & 15))
    fail("packed genotype rows must be 16-byte aligned with a pitch that is a multiple of 16 (pitch=%lld)", (long long)pitch);
  const int n_live = (pl.n_pad / 32 + 1) / 2 * 2, n_feat = n_wide + n_narrow;
  const int8_t* q = d_featq;
  for (int f0 = 0; f0 < n_wide; f0 += kTcWideFeatures, q += tc5_wide_chunk_bytes(pl.n_pad)) {
    e->prof_begin(kFamSeriesSums);   // one event pair per launch: bench.py's roofline divides by the launch count
    launch_tc5<kTcWideFeatures * 4, 512>(e, d_packed, pitch, n_snps, plink1, n_live, q, d_inv_scale, f0, n_feat, d_sums, nullptr, 0,
                                         nullptr, nullptr);
    e->prof_end(kFamSeriesSums);
    e->series_columns += kTcWideFeatures * 4;
  }
  for (int f0 = n_wide; f0 < n_feat; f0 += kTcNarrowFeatures, q += tc5_narrow_chunk_bytes(pl.n_pad)) {
    e->prof_begin(kFamSeriesSums);
    launch_tc5<kTcNarrowFeatures * 3, 512>(e, d_packed, pitch, n_snps, plink1, n_live, q, d_inv_scale, f0, n_feat, d_sums, nullptr, 0,
                                           nullptr, nullptr);
    e->prof_end(kFamSeriesSums);
    e->series_columns += kTcNarrowFeatures * 3;
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
    else launch_tc5<64, 512>(e, d_packed, pitch, n_snps, plink1, n_live, Bq, d_inv_scale, f0, n_feat, d_sums, d_cnt, count_slot, d_ml, d_mn);
  }
  e->prof_end(kFamClassSums);
}

}  // namespace gwsem
