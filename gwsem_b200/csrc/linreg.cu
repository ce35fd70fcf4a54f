// WLS marginals when a single continuous item is regressed on definition variables that include the
// SNP: buildItem(..., exogenous = TRUE) with a continuous depVar, the model of the reference's
// gene-environment vignette (tools/vignettes/GeneEnvironmentInteraction.Rmd:47-49; R/model.R:403-418,
// 427-434 for the `snp_<moderator>` product columns, 560-576 for the exogenous default).  The model is
// saturated: intercept, slopes and the ML residual variance of the regression are the parameters, and
// their asymptotic covariance in the outer-product form of Muthen's estimator, (sum_i sc_i sc_i')^-1
// with sc_i = (x_i e_i / s2, (e_i^2 - s2) / (2 s2^2)), is the parameter covariance
// (oracle/marginals_oracle.py, class Cont).
// One CTA per SNP, two passes over the people: X'X and X'y, then the residual moments.  Within a tile
// thread = person fills shared memory, then thread = (matrix entry, slice of the tile) accumulates.
#include "common.h"
#include "kernels.h"
#include "linreg.h"

namespace gwsem {

namespace {

constexpr int kT = 256;

template <bool DOSAGE>
__global__ void __launch_bounds__(kT)
    linreg_snp_kernel(LinregProgram lp, const uint8_t* __restrict__ rows, int64_t pitch, int n_src,
                      const uint8_t* __restrict__ inc8, int plink1, const double* __restrict__ geno, int n_pad,
                      const double* __restrict__ maf_in, gwsem_result res) {
  extern __shared__ double dyn[];
  const int snp = blockIdx.x, tid = threadIdx.x;
  const int d = lp.K + 1, dv = d + 1;              // regression coefficients; + residual variance
  const int n_tri = d * (d + 1) / 2;
  const int n_ent = n_tri + 2 * d;                 // pass 1: X'X, X'y (, unused); pass 2: Q, sum x e^3, sum x e
  const int slices = max(1, kT / n_ent);
  double* tX = dyn;                                // [kT][d]: regressors of the tile's people (zero rows if absent)
  double* tw = tX + kT * d;                        // [kT][3]: pass 1 (y, 0, 0); pass 2 (e^2, e^3, e)
  double* part = tw + kT * 3;                      // [slices][n_ent]
  double* G = part + slices * n_ent;               // d*d
  double* beta = G + d * d;                        // d
  double* M = beta + d;                            // dv*dv and its inverse (2 dv*dv)
  __shared__ double red[kT / 32];
  __shared__ int flag_sh;

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
  auto block_sum = [&](double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) s += red[w];
    return s;
  };
  const int my_ent = tid % n_ent, my_slice = tid / n_ent;
  const bool ent_active = my_slice < slices;
  int ea = 0, eb = -1, ekind = 0;   // entry (a, b) of the triangle, or vector entry a of kind 1 / 2
  if (my_ent < n_tri) {
    int e = my_ent;
    while (e > ea) { e -= ea + 1; ++ea; }
    eb = e;
  } else {
    ekind = 1 + (my_ent - n_tri) / d;
    ea = (my_ent - n_tri) % d;
  }
  const int per = (kT + slices - 1) / slices;
  const int p_lo = my_slice * per, p_hi = min(kT, p_lo + per);

  double cnt = 0.0, gsum = 0.0, gss = 0.0, yy = 0.0, s2sum = 0.0, s4sum = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    double acc = 0.0;
    for (int t0 = 0; t0 < n_src; t0 += kT) {
      const int i = t0 + tid;
      double g = 0.0;
      const bool live = i < n_src && genotype(i, &g);
      double* x = tX + tid * d;
      if (live) {
        x[0] = 1.0;
        const double* xs = lp.Xs + (int64_t)i * lp.Ksp;
        for (int k = 0; k < lp.K; ++k)
          x[1 + k] = lp.kind[k] == 0 ? xs[lp.col[k]] : (lp.kind[k] == 1 ? g : g * xs[lp.col[k]]);
        const double y = lp.y[i];
        if (pass == 0) {
          tw[tid * 3] = y; tw[tid * 3 + 1] = 0.0; tw[tid * 3 + 2] = 0.0;
          cnt += 1.0; gsum += g; gss = fma(g, g, gss); yy = fma(y, y, yy);
        } else {
          double e = y;
          for (int k = 0; k < d; ++k) e = fma(-beta[k], x[k], e);
          const double e2 = e * e;
          tw[tid * 3] = e2; tw[tid * 3 + 1] = e2 * e; tw[tid * 3 + 2] = e;
          s2sum += e2; s4sum = fma(e2, e2, s4sum);
        }
      } else {
        for (int k = 0; k < d; ++k) x[k] = 0.0;
        tw[tid * 3] = tw[tid * 3 + 1] = tw[tid * 3 + 2] = 0.0;
      }
      __syncthreads();
      if (ent_active) {
        for (int p = p_lo; p < p_hi; ++p) {
          const double xa = tX[p * d + ea];
          if (ekind == 0) acc = fma(xa * tX[p * d + eb], pass == 0 ? 1.0 : tw[p * 3], acc);
          else if (ekind == 1) acc = fma(xa, pass == 0 ? tw[p * 3] : tw[p * 3 + 1], acc);
          else acc = fma(xa, tw[p * 3 + 2], acc);
        }
      }
      __syncthreads();
    }
    if (ent_active) part[my_slice * n_ent + my_ent] = acc;
    __syncthreads();
    if (pass == 0) {
      const double n = block_sum(cnt), sg = block_sum(gsum), sgg = block_sum(gss);
      yy = block_sum(yy);
      cnt = n;
      if (tid < n_ent) {
        double v = 0.0;
        for (int s = 0; s < slices; ++s) v += part[s * n_ent + tid];
        if (ekind == 0) { G[ea * d + eb] = v; G[eb * d + ea] = v; }
        else if (ekind == 1) beta[ea] = v;           // X'y for now
      }
      __syncthreads();
      if (tid == 0) {
        // observed variance of the genotype (mxData minVariance), then beta = (X'X)^-1 X'y by Cholesky
        int state = 0;
        if (n < 2.0 || !((sgg - sg * sg / n) / (n - 1.0) >= lp.min_variance)) state = 1;
        for (int j = 0; j < d && !state; ++j) {
          double v = G[j * d + j];
          for (int k = 0; k < j; ++k) v -= G[j * d + k] * G[j * d + k];
          if (!(v > 1e-10 * G[j * d + j]) || !isfinite(v)) { state = 2; break; }
          G[j * d + j] = sqrt(v);
          for (int i2 = j + 1; i2 < d; ++i2) {
            double w = G[i2 * d + j];
            for (int k = 0; k < j; ++k) w -= G[i2 * d + k] * G[j * d + k];
            G[i2 * d + j] = w / G[j * d + j];
          }
        }
        if (!state) {
          for (int i2 = 0; i2 < d; ++i2) {
            double v = beta[i2];
            for (int k = 0; k < i2; ++k) v -= G[i2 * d + k] * beta[k];
            beta[i2] = v / G[i2 * d + i2];
          }
          for (int i2 = d - 1; i2 >= 0; --i2) {
            double v = beta[i2];
            for (int k = i2 + 1; k < d; ++k) v -= G[k * d + i2] * beta[k];
            beta[i2] = v / G[i2 * d + i2];
          }
        }
        flag_sh = state;
      }
      __syncthreads();
      if (flag_sh) break;
    }
  }
  // ---- results
  const int P = lp.P;
  double* est = res.est + (int64_t)snp * P;
  double* vcov = res.vcov + (int64_t)snp * P * P;
  const double na = __longlong_as_double((long long)GWSEM_NA_BITS);
  const int state = flag_sh;
  const double n = cnt;
  for (int k = tid; k < P * P; k += kT) vcov[k] = na;
  for (int k = tid; k < P; k += kT) est[k] = lp.par_start[k];
  if (tid == 0) {
    res.maf[snp] = maf_in[snp];
    res.n_used[snp] = (int)n;
    res.iterations[snp] = state ? 0 : 1;
    res.catch_var[snp] = state == 1 ? -2 : -1;
    res.catch_code[snp] = state == 1 ? GWSEM_CATCH_MIN_VARIANCE : (state == 2 ? GWSEM_CATCH_ACOV_NOT_PD : GWSEM_CATCH_NONE);
    res.status[snp] = state ? GWSEM_STATUS_NA : GWSEM_STATUS_OK;
    res.fit[snp] = state ? na : 0.0;   // saturated: the WLS discrepancy is zero
  }
  if (state) return;
  const double s2tot = block_sum(s2sum), s4tot = block_sum(s4sum);
  const double s2 = s2tot / n;
  // outer products of the scores: [[Q / s2^2, (sum x e^3 - s2 sum x e) / (2 s2^3)], [., (s4 - 2 s2 S2 + n s2^2) / (4 s2^4)]]
  if (tid < n_ent) {
    double v = 0.0;
    for (int s = 0; s < slices; ++s) v += part[s * n_ent + tid];
    part[tid] = v;   // slice 0 now holds the totals
  }
  __syncthreads();
  if (tid == 0) {
    const double i2 = 1.0 / (s2 * s2), i3 = i2 / s2, i4 = i2 * i2;
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b <= a; ++b) {
        const double v = part[a * (a + 1) / 2 + b] * i2;
        M[a * dv + b] = v; M[b * dv + a] = v;
      }
      const double c = (part[n_tri + a] - s2 * part[n_tri + d + a]) * 0.5 * i3;
      M[a * dv + d] = c; M[d * dv + a] = c;
    }
    M[d * dv + d] = (s4tot - 2.0 * s2 * s2tot + n * s2 * s2) * 0.25 * i4;
    // inverse by Gauss-Jordan (symmetric positive definite)
    double* V = M + dv * dv;
    bool ok = true;
    for (int k = 0; k < dv * dv; ++k) V[k] = (k / dv == k % dv) ? 1.0 : 0.0;
    for (int c = 0; c < dv && ok; ++c) {
      const double piv = M[c * dv + c];
      if (!(piv > 0.0) || !isfinite(piv)) { ok = false; break; }
      const double inv = 1.0 / piv;
      for (int k = 0; k < dv; ++k) { M[c * dv + k] *= inv; V[c * dv + k] *= inv; }
      for (int r = 0; r < dv; ++r) {
        if (r == c) continue;
        const double f = M[r * dv + c];
        if (f == 0.0) continue;
        for (int k = 0; k < dv; ++k) { M[r * dv + k] -= f * M[c * dv + k]; V[r * dv + k] -= f * V[c * dv + k]; }
      }
    }
    flag_sh = ok ? 1 : 0;
  }
  __syncthreads();
  for (int k = tid; k < P; k += kT) {
    const int m = lp.par_map[k];
    est[k] = m < d ? beta[m] : s2;
  }
  if (flag_sh) {
    const double* V = M + dv * dv;
    for (int k = tid; k < P * P; k += kT) vcov[k] = V[lp.par_map[k / P] * dv + lp.par_map[k % P]];
  } else if (tid == 0) {
    res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
  }
}

}  // namespace

void launch_linreg_snp(gwsem_engine* e, const LinregProgram& lp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, int plink1, const double* d_geno, int n_pad, const double* d_maf,
                       const gwsem_result& d_res) {
  if (n_snps <= 0) return;
  const int d = lp.K + 1, dv = d + 1, n_ent = d * (d + 1) / 2 + 2 * d;
  if (n_ent > kT) fail("regression with %d definition variables is more than this kernel handles", lp.K);
  const int slices = std::max(1, kT / n_ent);
  const size_t doubles = (size_t)kT * d + (size_t)kT * 3 + (size_t)slices * n_ent + (size_t)d * d + d + 2 * (size_t)dv * dv;
  const size_t smem = doubles * sizeof(double) + 64;
  e->prof_begin(kFamClassSums);
  if (d_geno) {
    GW_CUDA(cudaFuncSetAttribute(linreg_snp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linreg_snp_kernel<true><<<n_snps, kT, smem, e->stream>>>(lp, d_rows, pitch, n_src, d_inc8, plink1, d_geno, n_pad, d_maf, d_res);
  } else {
    GW_CUDA(cudaFuncSetAttribute(linreg_snp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linreg_snp_kernel<false><<<n_snps, kT, smem, e->stream>>>(lp, d_rows, pitch, n_src, d_inc8, plink1, d_geno, n_pad, d_maf, d_res);
  }
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamClassSums);
  e->launches++;
}

}  // namespace gwsem
