// WLS marginals when the SNP is an exogenous predictor of one ordinal item: buildItem() with an
// ordered-factor phenotype (reference R/model.R:560-576: exogenous defaults to TRUE, `snp` and the
// covariates become definition-variable latents with free slopes; README / man example
// buildItem(pheno, 'anxiety')).  The model is saturated and the item's variance is fixed at 1, so
// the summary statistics (thresholds, slopes of the ordered probit of the item on snp + covariates)
// are the parameter estimates, and their asymptotic covariance in the outer-product form of
// Muthen's estimator, (sum_i sc_i sc_i')^-1, is the parameter covariance.
// Per SNP: Newton-Raphson with the exact Hessian from the no-SNP fit; one CTA per SNP.
// Phase 1 (thread = person): Phi / phi of the two category bounds; phase 2 (thread = matrix entry):
// score, exact Hessian and outer-product information.
#include "common.h"
#include "kernels.h"
#include "probit.h"

namespace gwsem {

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double dphi(double z) { return 0.3989422804014327 * exp(-0.5 * z * z); }

template <bool DOSAGE>
__global__ void __launch_bounds__(kT)
    probit_snp_kernel(ProbitProgram pp, const uint8_t* __restrict__ rows, int64_t pitch, int n_src,
                      const uint8_t* __restrict__ inc8, int plink1, const double* __restrict__ geno, int n_pad,
                      const double* __restrict__ maf_in, gwsem_result res) {
  extern __shared__ double dyn[];
  const int snp = blockIdx.x, tid = threadIdx.x;
  const int C = pp.n_cat, K = pp.K, d = C - 1 + K;  // thresholds C-1, then a slope per definition variable
  const int n_ent = d * (d + 1) / 2 + d;           // lower triangle of H / OPG, then the score
  const int slices = max(1, kT / n_ent);
  double* theta = dyn;                 // d
  double* gsc = theta + d;             // d
  double* H = gsc + d;                 // d*d
  double* G = H + d * d;               // d*d (outer products)
  double* tA = G + d * d;              // [kT] p1/pi
  double* tB = tA + kT;                // p0/pi
  double* tD1 = tB + kT;               // -z1 p1 / pi
  double* tD0 = tD1 + kT;              // -z0 p0 / pi
  double* tg = tD0 + kT;               // genotype value
  double* tX = tg + kT;                // [kT][K]
  double* part = tX + kT * max(K, 1);  // [slices][n_ent][2]
  int* tcat = reinterpret_cast<int*>(part + (size_t)slices * n_ent * 2);  // [kT]
  __shared__ int n_sh, flag_sh;
  __shared__ double red_n[kT / 32];

  const uint8_t* row = DOSAGE ? nullptr : rows + (int64_t)snp * pitch;
  const double* grow = DOSAGE ? geno + (int64_t)snp * n_pad : nullptr;
  auto genotype = [&](int i, double* g) -> bool {
    if (!inc8[i]) return false;
    if (DOSAGE) { *g = grow[i]; return *g == *g; }
    unsigned code = (row[i >> 2] >> (2 * (i & 3))) & 3u;
    if (plink1) code = (0x1Eu >> (2 * code)) & 3u;
    *g = (double)code;
    return code != 3u;
  };
  // start from the fit without the SNP (slope 0)
  for (int k = tid; k < d; k += kT) theta[k] = pp.theta0[k];
  if (tid == 0) { n_sh = 0; flag_sh = 0; }
  __syncthreads();
  const int my_ent = tid % n_ent, my_slice = tid / n_ent;
  const bool ent_active = my_slice < slices;
  // decode entry -> (a, b); score entries have b = -1
  int ea = 0, eb = -1;
  if (my_ent < d * (d + 1) / 2) {
    int e = my_ent;
    while (e > ea) { e -= ea + 1; ++ea; }
    eb = e;  // 0 <= eb <= ea
  } else {
    ea = my_ent - d * (d + 1) / 2;
  }
  int status = GWSEM_STATUS_ITERATION_LIMIT, iters = 0;
  bool failed = false;
  // observed variance of the genotype among the people used (mxData minVariance)
  {
    double c0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < n_src; i += kT) {
      double g;
      if (genotype(i, &g)) { c0 += 1.0; s1 += g; s2 = fma(g, g, s2); }
    }
    __shared__ double mv[3][kT / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((tid & 31) == 0) { mv[0][tid >> 5] = c0; mv[1][tid >> 5] = s1; mv[2][tid >> 5] = s2; }
    __syncthreads();
    c0 = s1 = s2 = 0.0;
    for (int w = 0; w < kT / 32; ++w) { c0 += mv[0][w]; s1 += mv[1][w]; s2 += mv[2][w]; }
    if (c0 < 2.0 || !((s2 - s1 * s1 / c0) / (c0 - 1.0) >= pp.min_variance)) {
      const int P0 = pp.P;
      const double na0 = __longlong_as_double((long long)GWSEM_NA_BITS);
      for (int k = tid; k < P0 * P0; k += kT) res.vcov[(int64_t)snp * P0 * P0 + k] = na0;
      for (int k = tid; k < P0; k += kT) res.est[(int64_t)snp * P0 + k] = pp.par_start[k];
      if (tid == 0) {
        res.maf[snp] = maf_in[snp];
        res.n_used[snp] = (int)c0;
        res.iterations[snp] = 0;
        res.catch_var[snp] = -2;
        res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE;
        res.status[snp] = GWSEM_STATUS_NA;
        res.fit[snp] = na0;
      }
      return;
    }
  }
  for (int iter = 0; iter < 40; ++iter) {
    double accH = 0.0, accG = 0.0;
    int cnt = 0;
    for (int t0 = 0; t0 < n_src; t0 += kT) {
      const int i = t0 + tid;
      double A = 0.0, B = 0.0, D1 = 0.0, D0 = 0.0, g = 0.0;
      int cat = 0;
      if (i < n_src && genotype(i, &g)) {
        cat = pp.cat[i];
        double eta = 0.0;
        const double* xs = pp.X + (int64_t)i * pp.Kp;
        for (int k = 0; k < K; ++k) {
          const double x = pp.kind[k] == 0 ? xs[pp.col[k]] : (pp.kind[k] == 1 ? g : g * xs[pp.col[k]]);
          tX[tid * K + k] = x;
          eta = fma(theta[C - 1 + k], x, eta);
        }
        const double z1 = (cat < C - 1 ? theta[cat] : 12.0) - eta, z0 = (cat > 0 ? theta[cat - 1] : -12.0) - eta;
        const double p1 = cat < C - 1 ? dphi(z1) : 0.0, p0 = cat > 0 ? dphi(z0) : 0.0;
        const double pi = fmax(z0 > 0.0 ? normcdf(-z0) - normcdf(-z1) : normcdf(z1) - normcdf(z0), 1e-300);
        A = p1 / pi; B = p0 / pi; D1 = -z1 * A; D0 = -z0 * B;
        ++cnt;
      } else {
        for (int k = 0; k < K; ++k) tX[tid * K + k] = 0.0;
      }
      tA[tid] = A; tB[tid] = B; tD1[tid] = D1; tD0[tid] = D0; tg[tid] = g; tcat[tid] = cat;
      __syncthreads();
      if (ent_active) {
        const int per = (kT + slices - 1) / slices;
        const int p_lo = my_slice * per, p_hi = min(kT, p_lo + per);
        for (int p = p_lo; p < p_hi; ++p) {
          const double pA = tA[p], pB = tB[p];
          if (pA == 0.0 && pB == 0.0) continue;
          const int c = tcat[p];
          // derivatives of z1, z0 with respect to parameter a (threshold / slope)
          auto jz = [&](int a, double* j1, double* j0) {
            if (a < C - 1) { *j1 = (c == a) ? 1.0 : 0.0; *j0 = (c == a + 1) ? 1.0 : 0.0; }
            else { const double x = tX[p * K + (a - (C - 1))]; *j1 = -x; *j0 = -x; }
          };
          double a1, a0;
          jz(ea, &a1, &a0);
          const double ga = pA * a1 - pB * a0;
          if (eb < 0) { accG += ga; continue; }
          double b1, b0;
          jz(eb, &b1, &b0);
          const double gb = pA * b1 - pB * b0;
          accG = fma(ga, gb, accG);
          accH += ga * gb - (tD1[p] * a1 * b1 - tD0[p] * a0 * b0);
        }
      }
      __syncthreads();
    }
    // reduce the slices
    if (ent_active) { part[((size_t)my_slice * n_ent + my_ent) * 2] = accH; part[((size_t)my_slice * n_ent + my_ent) * 2 + 1] = accG; }
    if (iter == 0) {
      int c32 = cnt;
      for (int o = 16; o; o >>= 1) c32 += __shfl_xor_sync(0xffffffffu, c32, o);
      if ((tid & 31) == 0) atomicAdd(&n_sh, c32);
    }
    __syncthreads();
    if (tid < n_ent) {
      double h = 0.0, g = 0.0;
      for (int s = 0; s < slices; ++s) { h += part[((size_t)s * n_ent + tid) * 2]; g += part[((size_t)s * n_ent + tid) * 2 + 1]; }
      if (eb >= 0) { H[ea * d + eb] = h; H[eb * d + ea] = h; G[ea * d + eb] = g; G[eb * d + ea] = g; }
      else gsc[ea] = g;
    }
    __syncthreads();
    // Newton step by thread 0: solve H step = score (Cholesky)
    if (tid == 0) {
      bool ok = true;
      double* L = part;  // scratch
      for (int j = 0; j < d && ok; ++j) {
        double v = H[j * d + j];
        for (int k = 0; k < j; ++k) v -= L[j * d + k] * L[j * d + k];
        if (!(v > 1e-12 * fabs(H[j * d + j])) || !isfinite(v)) { ok = false; break; }
        L[j * d + j] = sqrt(v);
        for (int i2 = j + 1; i2 < d; ++i2) {
          double w = H[i2 * d + j];
          for (int k = 0; k < j; ++k) w -= L[i2 * d + k] * L[j * d + k];
          L[i2 * d + j] = w / L[j * d + j];
        }
      }
      double worst = 0.0;
      if (ok) {
        double* st = L + d * d;
        for (int i2 = 0; i2 < d; ++i2) {
          double v = gsc[i2];
          for (int k = 0; k < i2; ++k) v -= L[i2 * d + k] * st[k];
          st[i2] = v / L[i2 * d + i2];
        }
        for (int i2 = d - 1; i2 >= 0; --i2) {
          double v = st[i2];
          for (int k = i2 + 1; k < d; ++k) v -= L[k * d + i2] * st[k];
          st[i2] = v / L[i2 * d + i2];
        }
        for (int k = 0; k < d; ++k) { theta[k] += st[k]; worst = fmax(worst, fabs(st[k])); }
      }
      flag_sh = !ok ? 2 : (worst < 1e-9 ? 1 : 0);
    }
    __syncthreads();
    iters = iter + 1;
    if (flag_sh == 2) { failed = true; break; }
    if (flag_sh == 1) { status = GWSEM_STATUS_OK; break; }
  }
  // ---- results: parameters are the statistics; vcov = (outer products)^-1 at the last evaluation
  const int P = pp.P;
  double* est = res.est + (int64_t)snp * P;
  double* vcov = res.vcov + (int64_t)snp * P * P;
  const double na = __longlong_as_double((long long)GWSEM_NA_BITS);
  for (int k = tid; k < P * P; k += kT) vcov[k] = na;
  if (tid == 0) {
    res.maf[snp] = maf_in[snp];
    res.n_used[snp] = n_sh;
    res.iterations[snp] = iters;
    res.catch_var[snp] = -1;
    res.catch_code[snp] = failed ? GWSEM_CATCH_ACOV_NOT_PD : GWSEM_CATCH_NONE;
    res.status[snp] = failed ? GWSEM_STATUS_NA : status;
    res.fit[snp] = failed ? na : 0.0;   // saturated: the WLS discrepancy is zero
  }
  for (int k = tid; k < P; k += kT) est[k] = failed ? pp.par_start[k] : theta[pp.par_map[k]];
  __syncthreads();
  if (failed) return;
  if (tid == 0) {  // invert G in place (Gauss-Jordan on the SPD matrix)
    bool ok = true;
    double* V = part;
    for (int k = 0; k < d * d; ++k) V[k] = (k / d == k % d) ? 1.0 : 0.0;
    for (int c = 0; c < d && ok; ++c) {
      const double piv = G[c * d + c];
      if (!(fabs(piv) > 0.0) || !isfinite(piv)) { ok = false; break; }
      const double inv = 1.0 / piv;
      for (int k = 0; k < d; ++k) { G[c * d + k] *= inv; V[c * d + k] *= inv; }
      for (int r = 0; r < d; ++r) {
        if (r == c) continue;
        const double f = G[r * d + c];
        if (f == 0.0) continue;
        for (int k = 0; k < d; ++k) { G[r * d + k] -= f * G[c * d + k]; V[r * d + k] -= f * V[c * d + k]; }
      }
    }
    flag_sh = ok ? 1 : 0;
  }
  __syncthreads();
  if (flag_sh)
    for (int k = tid; k < P * P; k += kT) vcov[k] = part[pp.par_map[k / P] * d + pp.par_map[k % P]];
  else if (tid == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
}

}  // namespace

void launch_probit_snp(gwsem_engine* e, const ProbitProgram& pp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, int plink1, const double* d_geno, int n_pad, const double* d_maf,
                       const gwsem_result& d_res) {
  if (n_snps <= 0) return;
  const int d = pp.n_cat - 1 + pp.K, n_ent = d * (d + 1) / 2 + d;
  if (n_ent > kT) fail("probit: %d parameters per SNP is more than this kernel handles", d);
  const int slices = std::max(1, kT / n_ent);
  const size_t doubles = (size_t)2 * d + 2 * d * d + 5 * kT + (size_t)kT * std::max(pp.K, 1) +
                         std::max<size_t>((size_t)slices * n_ent * 2, (size_t)d * d + d);
  const size_t smem = doubles * sizeof(double) + kT * sizeof(int) + 64;
  e->prof_begin(kFamClassSums);
  if (d_geno) {
    GW_CUDA(cudaFuncSetAttribute(probit_snp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probit_snp_kernel<true><<<n_snps, kT, smem, e->stream>>>(pp, d_rows, pitch, n_src, d_inc8, plink1, d_geno, n_pad, d_maf, d_res);
  } else {
    GW_CUDA(cudaFuncSetAttribute(probit_snp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probit_snp_kernel<false><<<n_snps, kT, smem, e->stream>>>(pp, d_rows, pitch, n_src, d_inc8, plink1, d_geno, n_pad, d_maf, d_res);
  }
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamClassSums);
  e->launches++;
}

}  // namespace gwsem
