// WLS "marginals": from the per-SNP summary record (marginals.cu) to Muthen's asymptotic
// covariance, the statistic vector, and the fit, one warp per SNP:
//   theta = (snp mean, snp var | item stage-1 | rho(snp,item_j) | rho(item,item))
//   acov(theta) = A^-1 B A^-T with A = [[A11, 0], [A21, A22]] in outer-product form
//   statistics s = (means | thresholds, slopes, variances, covariances) by the delta method;
//   a continuous variable contributes sd * rho to a covariance
//   model-implied vector: RAM moments conditional on the exogenous covariates, ordinal rows
//   standardised by the implied sd; damped Gauss-Newton with the analytic Jacobian.
// Mirrors oracle/marginals_oracle.py (fit_snp_marginals).  Two kernels (marg_fit_impl.cuh): marg_fit_kernel, the generic
// one (full records of marg_exact, any model), and marg_fit_fast_kernel, the same estimator with the SNP-independent
// algebra taken out of the loop for records assembled from the SNP-independent item side (series path / marg_stats).
#include <algorithm>

#include "common.h"
#include "fit_device.cuh"
#include "kernels.h"
#include "marginals.h"

namespace gwsem {

namespace {

// The RAM evaluation (ModelEval) needs only these pieces of the cumulants kernel's scratch layout.
__host__ __device__ inline WarpScratch marg_ws(int p, int nv) {
  WarpScratch w{};
  int o = 0;
  auto take = [&](int n) { int at = o; o += n; return at; };
  w.A = take(nv * nv); w.S = take(nv * nv); w.Z = take(nv * nv); w.C = take(nv * nv);
  w.T = take(nv * nv > 2 * nv ? nv * nv : 2 * nv);
  w.sig = take(p * (p + 1) / 2);
  w.total = o;
  return w;
}

struct MargScratch {
  int th, A, B, Ai, T, s, hidx, hcoef, L, Lb, scale, theta, cand, sig, sigp, r, rc, J, g, H, Hc, delta, total;
};

__host__ __device__ inline MargScratch marg_layout(int nt, int ns, int P, int nb_sep) {
  MargScratch m;
  int o = 0;
  auto take = [&](int n) { int at = o; o += n; return at; };
  // Four large regions, each used twice: first for Muthen's A^-1 B A^-T, then by the fit.
  //   A  -> acov -> Hc      B -> L (factor of the statistics' covariance)
  //   Ai -> J               T -> H
  auto big = [](int a, int b) { return a > b ? a : b; };
  m.th = take(nt);
  m.A = m.Hc = take(big(nt * nt, P * P));
  m.B = m.L = take(big(nt * nt, ns * ns));
  m.Ai = m.J = take(big(nt * nt, ns * P));
  m.T = m.H = take(big(nt * nt, P * P));
  m.s = take(ns);
  m.hidx = take(3 * ns);   // stored as doubles for simplicity
  m.hcoef = take(3 * ns);
  m.scale = take(ns);
  m.theta = take(P);
  m.cand = take(P);
  m.sig = take(ns);
  m.sigp = take(ns);
  m.r = take(ns);
  m.rc = take(ns);
  m.g = take(P);
  m.delta = take(P);
  m.Lb = take(nb_sep * nb_sep);  // factor of the b-block of the statistics' covariance (when rows are eliminated)
  m.total = o;
  return m;
}

}  // namespace

#define GW_MF_LANES 32
#define GW_MF_NS mf32
#include "marg_fit_impl.cuh"
#undef GW_MF_LANES
#undef GW_MF_NS
#define GW_MF_LANES 128
#define GW_MF_NS mf128
#include "marg_fit_impl.cuh"
#undef GW_MF_LANES
#undef GW_MF_NS

void launch_marg_pattern(gwsem_engine* e, const FitProgram& fp, const MargProgram& mp, double* d_out) {
  const int nt = mp.nt;
  const size_t smem = ((size_t)marg_ws(fp.p, fp.nv).total + marg_layout(nt, mp.n_stat, fp.P, 0).total) * sizeof(double);
  if (smem > 227 * 1024) fail("marginals model too large for the per-SNP fit kernel (%zu bytes of scratch)", smem);
  GW_CUDA(cudaFuncSetAttribute(mf32::marg_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mf32::marg_pattern_kernel<<<1, 32, smem, e->stream>>>(fp, mp, d_out);
  GW_CUDA(cudaGetLastError());
  e->launches++;
}

size_t marg_fit_scratch_doubles(const FitProgram& fp, const MargProgram& mp) {
  const int nt = mp.nt;
  return (size_t)marg_ws(fp.p, fp.nv).total + marg_layout(nt, mp.n_stat, fp.P, mp.n_a > 0 ? mp.n_b : 0).total;
}

void launch_marg_fit(gwsem_engine* e, const FitProgram& fp, const MargProgram& mp, int n_snps, const double* d_summary,
                     const double* d_maf, const gwsem_result& d_res) {
  if (n_snps <= 0) return;
  const size_t per = marg_fit_scratch_doubles(fp, mp) * sizeof(double);
  // one CTA of 128 threads per SNP: the scratch sets how many SNPs an SM holds (5 at the C3 shape), four warps on
  // each of them keep its issue slots busier than one warp per SNP did
  if (per > 227 * 1024) fail("marginals model too large for the per-SNP fit kernel (%zu bytes of scratch)", per);
  const size_t smem = per;
  // Records assembled from the SNP-independent item side go through the structured one-warp kernel (GWSEM_FIT_GENERIC=1
  // keeps everything on the generic one, an A/B aid); full records (marg_exact) stay on the generic kernel.
  const bool generic_only = getenv("GWSEM_FIT_GENERIC") != nullptr;   // read per launch (tests compare both in one process)
  MargProgram mpl = mp;
  mpl.fast_fit = 0;
  if (!generic_only && !mp.full_all && d_summary && mp.n_a > 0 && mp.Wcc) {
    const int D = 2 + mp.J;
    const size_t fper = ((size_t)marg_ws(fp.p, fp.nv).total + mf32::fast_layout(mp.nt, D, mp.q0, mp.n_stat, fp.P, mp.n_a, mp.n_b, mp.n_beta).total) * sizeof(double);
    const size_t tables = (size_t)mf32::fast_table_bytes(fp, mp);
    static const int warps_env = getenv("GWSEM_FIT_WARPS") ? atoi(getenv("GWSEM_FIT_WARPS")) : 0;   // tuning aid
    int warps = warps_env > 0 ? warps_env : 8;
    while (warps > 1 && warps * fper + tables > 225 * 1024) --warps;
    if (warps * fper + tables <= 225 * 1024) {
      const size_t fsmem = warps * fper + tables;
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (225 * 1024) / (fsmem + 2048)));
      const int grid = std::min((n_snps + warps - 1) / warps, e->sm_count * per_sm);
      mpl.fast_fit = 1;
      static const bool diag = getenv("GWSEM_FIT_DIAG") != nullptr;
      mpl.diag = diag;
      e->prof_begin(kFamFit);
      GW_CUDA(cudaFuncSetAttribute(mf32::marg_fit_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      mf32::marg_fit_fast_kernel<<<grid, 32 * warps, fsmem, e->stream>>>(fp, mpl, n_snps, d_summary, d_maf, d_res);
      GW_CUDA(cudaGetLastError());
      e->prof_end(kFamFit);
      e->launches++;
      if (diag) {
        unsigned long long h[8], z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        GW_CUDA(cudaStreamSynchronize(e->stream));
        GW_CUDA(cudaMemcpyFromSymbol(h, mf32::g_fast_clk, sizeof(h)));
        GW_CUDA(cudaMemcpyToSymbol(mf32::g_fast_clk, z, sizeof(z)));
        fprintf(stderr, "marg_fit_fast: %d SNPs, mean kcycles acov rows %.1f, gamma %.1f, iterations %.1f (%.2f), eliminated %.1f, vcov %.1f; %zu B of scratch\n",
                n_snps, h[0] / 1e3 / n_snps, h[1] / 1e3 / n_snps, h[2] / 1e3 / n_snps, (double)h[6] / n_snps, h[3] / 1e3 / n_snps, h[4] / 1e3 / n_snps, fper);
      }
    }
  }
  static const bool one_warp = getenv("GWSEM_FIT_ONE_WARP") != nullptr;   // A/B aid: the one-warp-per-SNP kernel
  e->prof_begin(kFamFit);
  if (one_warp) {
    GW_CUDA(cudaFuncSetAttribute(mf32::marg_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mf32::marg_fit_kernel<<<n_snps, 32, smem, e->stream>>>(fp, mpl, n_snps, d_summary, d_maf, d_res);
  } else {
    GW_CUDA(cudaFuncSetAttribute(mf128::marg_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mf128::marg_fit_kernel<<<n_snps, 128, smem, e->stream>>>(fp, mpl, n_snps, d_summary, d_maf, d_res);
  }
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamFit);
  e->launches++;
}

}  // namespace gwsem
