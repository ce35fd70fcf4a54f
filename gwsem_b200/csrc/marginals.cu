// Dispatcher of the per-SNP marginals statistics kernel (marg_stats_impl.cuh; instantiated per item
// count in marg_stats_j*.cu) and the MAF-from-counts kernel.
#include "common.h"
#include "kernels.h"
#include "marginals.h"

namespace gwsem {

template <int JJ>
void launch_marg_stats_j(gwsem_engine* e, const MargProgram& mp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                         int n_src, const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno,
                         int n_pad, double* d_summary, size_t smem);
#define GW_DECL(JJ)                                                                                                  \
  extern template void launch_marg_stats_j<JJ>(gwsem_engine*, const MargProgram&, const uint8_t*, int64_t, int, int, \
                                               const uint8_t*, const int32_t*, int, const double*, int, double*, size_t);
GW_DECL(1) GW_DECL(2) GW_DECL(3) GW_DECL(4) GW_DECL(5) GW_DECL(6) GW_DECL(7) GW_DECL(8) GW_DECL(9) GW_DECL(10)
#undef GW_DECL

void launch_marg_stats(gwsem_engine* e, const MargProgram& mp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno,
                       int n_pad, double* d_summary) {
  if (n_snps <= 0) return;
  constexpr int kT = 256;
  const int J = mp.J, D = 2 + J;
  const size_t smem = sizeof(double) * std::max<size_t>(std::max<size_t>((size_t)kT * (((D + 1) & ~1) + (mp.n1 | 1)) + 64, (size_t)8 * 64 * (D + 1)),
                                                         (size_t)8 * 64 + 4096 / 2);
  e->prof_begin(kFamMargStats);
#define GW_MARG(JJ) case JJ: launch_marg_stats_j<JJ>(e, mp, d_rows, pitch, n_snps, n_src, d_inc8, d_counts, plink1, d_geno, n_pad, d_summary, smem); break;
  switch (J) {
    GW_MARG(1) GW_MARG(2) GW_MARG(3) GW_MARG(4) GW_MARG(5) GW_MARG(6) GW_MARG(7) GW_MARG(8) GW_MARG(9) GW_MARG(10)
    default: fail("marginals: %d items (1..%d supported)", J, kMargMaxItems);
  }
#undef GW_MARG
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamMargStats);
  e->launches++;
}

// MAF from the class counts over the people kept by rowFilter (loader.cpp:386-394)
__global__ void maf_from_counts_kernel(const int32_t* __restrict__ counts, int n_kept, int n_snps, double* __restrict__ maf) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_snps) return;
  const int32_t* c = counts + 8 * (int64_t)r;
  double m = ((double)c[3] + 2.0 * (double)c[4]) / (2.0 * (double)(n_kept - c[5]));
  if (m > 0.5) m = 1 - m;
  maf[r] = m;
}

void launch_maf_from_counts(gwsem_engine* e, const int32_t* d_counts, int n_kept, int n_snps, double* d_maf) {
  if (n_snps <= 0) return;
  maf_from_counts_kernel<<<(n_snps + 255) / 256, 256, 0, e->stream>>>(d_counts, n_kept, n_snps, d_maf);
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

}  // namespace gwsem
