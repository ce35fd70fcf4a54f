// Per-SNP half of the WLS "marginals" summary statistics for models
// whose manifest variables are `snp` (continuous, no covariates: exoFree row all FALSE,
// reference R/model.R:413-416) and items (ordinal or continuous, regressed on the exogenous
// covariates).  Per SNP, with the item-side quantities precomputed by marginals_host.cpp:
//   stage 1 of snp            mean, ML variance (from the genotype class counts)
//   stage 2 snp x item j      polyserial (ordinal item) / Pearson (continuous item) correlation by
//                             quasi-Newton on the score with the outer-product information
//   Muthen's Gamma pieces     B rows of the SNP-dependent scores against every score column,
//                             A21 rows of the new correlations, A22
// One CTA per SNP.  FP64 throughout; Phi through CUDA's normcdf.
#pragma once
#include "common.h"
#include "kernels.h"
#include "marginals.h"

namespace gwsem {

namespace {

constexpr int kT = 256;  // threads per CTA = people per tile

__device__ __forceinline__ double dphi(double z) { return 0.3989422804014327 * exp(-0.5 * z * z); }
__device__ __forceinline__ double rect_prob(double a1, double a0) {
  // Phi(a1) - Phi(a0), computed in the tail where both are small
  return a0 > 0.0 ? normcdf(-a0) - normcdf(-a1) : normcdf(a1) - normcdf(a0);
}

// 1/x to full double precision for normal x: 20-bit hardware estimate + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}
// exp(x): Taylor to x^9 when |x| <= 1/4 (truncation below 3e-13 relative), else the library call
__device__ __forceinline__ double exp_small(double x) {
  if (fabs(x) > 0.25) return exp(x);
  double p = 1.0 / 362880.0;
  p = fma(p, x, 1.0 / 40320.0);
  p = fma(p, x, 1.0 / 5040.0);
  p = fma(p, x, 1.0 / 720.0);
  p = fma(p, x, 1.0 / 120.0);
  p = fma(p, x, 1.0 / 24.0);
  p = fma(p, x, 1.0 / 6.0);
  p = fma(p, x, 0.5);
  p = fma(p, x, 1.0);
  return fma(p, x, 1.0);
}

// The same two series two orders shorter, for SNPs whose shifts are small (the usual case: d ~ rho u with
// rho = O(1/sqrt(n))).  marg_stats_kernel picks them per SNP and item from a bound on the shifts:
// |d| <= kLowD leaves a truncation error below |He4(z)| phi(z) d^5 / 120 < 2.5e-10 in Phi, and
// |x| <= kLowX one below x^7 / 5040 < 1e-10 in exp.
constexpr double kLowD = 0.03, kLowX = 0.125;
__device__ __forceinline__ double exp_small_low(double x) {
  double p = 1.0 / 720.0;
  p = fma(p, x, 1.0 / 120.0);
  p = fma(p, x, 1.0 / 24.0);
  p = fma(p, x, 1.0 / 6.0);
  p = fma(p, x, 0.5);
  p = fma(p, x, 1.0);
  return fma(p, x, 1.0);
}
__device__ __forceinline__ double phi_integral_poly_low(double z, double d) {
  const double z2 = z * z;
  double p = (z2 * z - 3.0 * z) * (1.0 / 24.0);
  p = fma(p, d, (z2 - 1.0) * (1.0 / 6.0));
  p = fma(p, d, z * 0.5);
  p = fma(p, d, 1.0);
  return p * d;
}

struct PairTerms {
  double psi;      // d loglik_i / d rho
  double dpsi;     // d psi / d rho (exact, for Newton)
  double dl_du;    // d loglik2_i / d u (the standardised genotype)
  double dl_a;     // ordinal item: d/dz1 ; continuous item: d/dv
  double dl_b;     // ordinal item: d/dz0
};

// Phi(z) - Phi(z - d) = phi(z) * P(z, d): Taylor in d with Hermite polynomials, |d| <= kTaylorMax
constexpr double kTaylorMax = 0.125;
__device__ __forceinline__ double phi_integral_poly(double z, double d) {
  const double z2 = z * z;
  double p = (z2 * z2 * z - 10.0 * z2 * z + 15.0 * z) * (1.0 / 720.0);
  p = fma(p, d, (z2 * z2 - 6.0 * z2 + 3.0) * (1.0 / 120.0));
  p = fma(p, d, (z2 * z - 3.0 * z) * (1.0 / 24.0));
  p = fma(p, d, (z2 - 1.0) * (1.0 / 6.0));
  p = fma(p, d, z * 0.5);
  p = fma(p, d, 1.0);
  return p * d;
}

// zz = {z1, z0, phi(z1), phi(z0), Phi(z1) - Phi(z0)} of the person for an ordinal item (the +-12
// sentinels of the open categories make their phi vanish), {v, ...} for a continuous item.
// The rho-dependent scalars come from shared memory (pc = {rho, 1/sqrt(1-rho^2), its cube, 3 rho / (1-rho^2)}):
// kept in registers the compiler re-derived them, rsqrt included, for every person.
__device__ __forceinline__ PairTerms pair_terms(bool ordinal, const double* pc, double u, const double* zz, bool low = false) {
  PairTerms t;
  const double rho = pc[0], inv_s = pc[1];
  if (ordinal) {
    const double z1 = zz[0], z0 = zz[1];
    const double a1 = (z1 - rho * u) * inv_s, a0 = (z0 - rho * u) * inv_s;
    const double d1 = z1 - a1, d0 = z0 - a0;
    double f1, f0, pi;
    if (low) {   // uniform over the CTA: every person of this SNP is inside the short series' range
      f1 = zz[2] * exp_small_low(zz[2] == 0.0 ? 0.0 : fma(z1, d1, -0.5 * d1 * d1));
      f0 = zz[3] * exp_small_low(zz[3] == 0.0 ? 0.0 : fma(z0, d0, -0.5 * d0 * d0));
      pi = zz[4] - zz[2] * phi_integral_poly_low(z1, d1) + zz[3] * phi_integral_poly_low(z0, d0);
    } else if (fabs(d1) <= kTaylorMax && fabs(d0) <= kTaylorMax) {
      // phi(z - d) = phi(z) exp(z d - d^2/2);  Phi(z - d) = Phi(z) - phi(z) P(z, d)
      // (an open category's bound carries phi = 0 exactly: keep its argument out of the slow path)
      f1 = zz[2] * exp_small(zz[2] == 0.0 ? 0.0 : fma(z1, d1, -0.5 * d1 * d1));
      f0 = zz[3] * exp_small(zz[3] == 0.0 ? 0.0 : fma(z0, d0, -0.5 * d0 * d0));
      pi = zz[4] - zz[2] * phi_integral_poly(z1, d1) + zz[3] * phi_integral_poly(z0, d0);
    } else {
      f1 = dphi(a1);
      f0 = dphi(a0);
      pi = rect_prob(a1, a0);
    }
    pi = fmax(pi, 1e-300);
    const double inv_pi = fast_rcp(pi), s3 = pc[2], r3 = pc[3];
    const double g1 = (rho * z1 - u) * s3, g0 = (rho * z0 - u) * s3;          // d a / d rho
    const double h1 = fma(r3, g1, z1 * s3), h0 = fma(r3, g0, z0 * s3);
    t.psi = (f1 * g1 - f0 * g0) * inv_pi;
    t.dpsi = (f1 * (h1 - a1 * g1 * g1) - f0 * (h0 - a0 * g0 * g0)) * inv_pi - t.psi * t.psi;
    t.dl_a = f1 * inv_s * inv_pi;
    t.dl_b = -f0 * inv_s * inv_pi;
    t.dl_du = -u - rho * inv_s * (f1 - f0) * inv_pi;
  } else {
    const double v = zz[0], s2 = 1.0 - rho * rho, is2 = 1.0 / s2;
    const double num = rho * s2 + u * v * (1.0 + rho * rho) - rho * (u * u + v * v);
    t.psi = num * is2 * is2;
    const double dnum = 1.0 - 3.0 * rho * rho + 2.0 * rho * u * v - (u * u + v * v);
    t.dpsi = dnum * is2 * is2 + 4.0 * rho * num * is2 * is2 * is2;
    t.dl_du = -(u - rho * v) * is2;
    t.dl_a = -(v - rho * u) * is2;
    t.dl_b = 0.0;
  }
  return t;
}

__device__ __forceinline__ double block_sum(double v, double* red, int tid) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kT / 32; ++w) s += red[w];
  return s;
}

// DOSAGE = false: hardcalls from the packed 2-bit row (class counts given);
// DOSAGE = true: genotype column of doubles in file order (NaN = missing), moments taken here.
template <int J, bool DOSAGE, int HALVES>
__global__ void __launch_bounds__(kT, 2)
    marg_stats_kernel(MargProgram mp, const uint8_t* __restrict__ rows, int64_t pitch, int n_src,
                      const uint8_t* __restrict__ inc8, const int32_t* __restrict__ counts, int plink1,
                      const double* __restrict__ geno, int n_pad, double* __restrict__ summary) {
  constexpr int D = 2 + J;
  extern __shared__ __align__(16) double dyn[];
  __shared__ double red[kT / 32];
  __shared__ double rho_sh[J];
  __shared__ double pc_sh[J][4];
  __shared__ int low_sh[J];
  __shared__ double sums_sh[J][3];
  __shared__ int done_sh;
  const int snp = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int se = (!DOSAGE && mp.se_done) ? mp.se_done[snp] : 0;
  if (se == 1) return;   // the series path (marg_series.cu) has written this record
  double* out = summary + (int64_t)snp * mp.sum_stride;
  // se == 2: the series path left the roots of its order-5 score polynomials in the record: correlations within this
  // kernel's own Newton stopping distance of the estimates (read before the record is cleared)
  double rho_given = 0.0;
  if (se == 2 && tid < J) rho_given = out[mp.o_rho + tid];
  __syncthreads();
  int n;
  double dn, mu, var;
  const double* grow = DOSAGE ? geno + (int64_t)snp * n_pad : nullptr;
  if (!DOSAGE) {
    const int32_t* c = counts + 8 * (int64_t)snp;
    n = mp.n_incl - c[2];
    dn = (double)n;
    mu = ((double)c[0] + 2.0 * (double)c[1]) / dn;
    const double n0 = dn - (double)c[0] - (double)c[1];
    var = (n0 * mu * mu + (double)c[0] * (1.0 - mu) * (1.0 - mu) + (double)c[1] * (2.0 - mu) * (2.0 - mu)) / dn;
  } else {
    double cnt = 0.0, sum = 0.0;
    for (int i = tid; i < n_src; i += kT) {
      const double g = grow[i];
      if (inc8[i] && g == g) { cnt += 1.0; sum += g; }
    }
    dn = block_sum(cnt, red, tid);
    mu = block_sum(sum, red, tid) / dn;
    double ss = 0.0;
    for (int i = tid; i < n_src; i += kT) {
      const double g = grow[i];
      if (inc8[i] && g == g) ss = fma(g - mu, g - mu, ss);
    }
    var = block_sum(ss, red, tid) / dn;
    n = (int)dn;
  }
  for (int k = tid; k < mp.sum_stride; k += kT) out[k] = 0.0;
  __syncthreads();
  if (tid == 0) { out[0] = dn; out[1] = mu; out[2] = var; }
  if (!(var > 0.0) || n < 2) return;
  const double sd = sqrt(var);
  const double uc[3] = {(0.0 - mu) / sd, (1.0 - mu) / sd, (2.0 - mu) / sd};
  const double umax = fmax(fabs(uc[0]), fabs(uc[2]));   // genotypes and dosages lie in [0, 2]
  const uint8_t* row = DOSAGE ? nullptr : rows + (int64_t)snp * pitch;
  // standardised genotype of person i; returns false for missing calls and excluded people
  auto std_geno = [&](int i, double* u) -> bool {
    if (!inc8[i]) return false;
    if (DOSAGE) {
      const double g = grow[i];
      *u = (g - mu) / sd;
      return g == g;
    }
    unsigned code = (row[i >> 2] >> (2 * (i & 3))) & 3u;
    if (plink1) code = (0x1Eu >> (2 * code)) & 3u;
    if (code == 3u) return false;
    *u = uc[code];
    return true;
  };
  bool ordinal[J];
#pragma unroll
  for (int j = 0; j < J; ++j) ordinal[j] = mp.n_cat[j] > 0;

  // A missing call drops the person from every statistic of this SNP (naAction = 'omit'), the items' own
  // included: nothing on the item side is SNP-independent any more, and marg_exact_kernel recomputes the lot.
  if (n < mp.n_incl) return;

  // ---- stage 2: Newton on the score of every rho_j (exact second derivative; the outer-product
  // information stands in when the curvature estimate is not negative)
  if (tid < J) rho_sh[tid] = se == 2 ? rho_given : 0.0;
  if (tid == 0) done_sh = 0;
  __syncthreads();
  // Iteration 0 is at rho = 0, where the tetrachoric series gives the score and its first two
  // derivatives without exp or Phi: with h_k = He_k(u) c_k and the person's
  // c_k = (He_{k-1}(z0) phi(z0) - He_{k-1}(z1) phi(z1)) / pi (marginals_host side, mp.cc),
  // psi = h1, psi' = h2 - h1^2, psi'' = h3 - 3 h1 h2 + 2 h1^3.  The root of that quadratic model is
  // within O(rho^3) of the estimate; when it is below kSeriesMax the loop ends there, otherwise
  // exact Newton passes follow until a step is below kNewtonStop.  The Gamma pass below evaluates
  // one more exact Newton step at no extra cost and applies it.
  // (the score rows of the Gamma pass are evaluated before that last step: its size bounds the relative error of
  // the weight matrix, and with it the estimates' distance from the oracle's in units of their standard errors)
  // The series root is O(rho^3) from the estimate: accepted below 0.012 (2e-6), so that the scores of the Gamma pass,
  // taken at the accepted point, stand within 1e-5 of an SE of those at the estimate.
  constexpr double kNewtonStop = 3e-5, kFinalClamp = 3e-4, kSeriesMax = 0.012;
  for (int iter = 0; iter < 30 && se != 2; ++iter) {
    double s1[J], s2[J], s3[J], rh[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { s1[j] = s2[j] = s3[j] = 0.0; rh[j] = rho_sh[j]; }
    if (tid < J) {
      const double r = rho_sh[tid], is = rsqrt(1.0 - r * r);
      pc_sh[tid][0] = r; pc_sh[tid][1] = is; pc_sh[tid][2] = is * is * is; pc_sh[tid][3] = 3.0 * r * is * is;
      const double dmax = mp.zmax[tid] * (is - 1.0) + fabs(r) * umax * is;
      low_sh[tid] = mp.n_cat[tid] > 0 && dmax <= kLowD && mp.zmax[tid] * dmax + 0.5 * dmax * dmax <= kLowX;
    }
    __syncthreads();
    // Iteration 0 without a pass over people: for hardcalls u takes three values, so the three sums are combinations
    // of genotype-class sums of six per-person products (mp.cs_sums, exact fixed point; class 0 = totals - het - hom).
    const bool from_sums = iter == 0 && !DOSAGE && mp.cs_sums != nullptr && n == mp.n_incl;
    if (from_sums) {
      if (tid < J) {
        const double* S1 = mp.cs_sums + ((int64_t)snp * 2 + 0) * mp.cs_nf + 6 * tid;
        const double* S2 = mp.cs_sums + ((int64_t)snp * 2 + 1) * mp.cs_nf + 6 * tid;
        const double* T = mp.cs_tot + 6 * tid;
        double a = 0.0, b = 0.0, c = 0.0;
        const int32_t* cn = counts + 8 * (int64_t)snp;
        const double ncls[3] = {dn - (double)cn[0] - (double)cn[1], (double)cn[0], (double)cn[1]};
        const bool ord = mp.n_cat[tid] > 0;
#pragma unroll
        for (int cls = 0; cls < 3; ++cls) {
          double S[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) S[k] = cls == 0 ? T[k] - S1[k] - S2[k] : (cls == 1 ? S1[k] : S2[k]);
          const double uu = uc[cls], he2 = uu * uu - 1.0, he3 = uu * (uu * uu - 3.0);
          a += uu * S[0];
          if (ord) {
            b += he2 * S[1] - uu * uu * S[2];
            c += he3 * S[3] + 2.0 * uu * uu * uu * S[4] - 3.0 * uu * he2 * S[5];
          } else {   // continuous item: S[0] = sum v, S[1] = sum v^2
            b += -he2 * ncls[cls] - S[1];
          }
        }
        if (!ord) c = 6.0 * a;
        sums_sh[tid][0] = a; sums_sh[tid][1] = b; sums_sh[tid][2] = c;
      }
      __syncthreads();
    } else if (iter == 0) {
      for (int i = tid; i < n_src; i += kT) {
        double u;
        if (!std_geno(i, &u)) continue;
        const double he2 = u * u - 1.0, he3 = u * (u * u - 3.0);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const double* cc = mp.cc + ((int64_t)i * J + j) * 3;
          double psi, dpsi, d2psi;
          if (ordinal[j]) {
            const double h1 = u * cc[0], h2 = he2 * cc[1], h3 = he3 * cc[2];
            psi = h1;
            dpsi = h2 - h1 * h1;
            d2psi = h3 + h1 * (2.0 * h1 * h1 - 3.0 * h2);
          } else {
            const double v = cc[0];
            psi = u * v;
            dpsi = 1.0 - (u * u + v * v);
            d2psi = 6.0 * psi;
          }
          s1[j] += psi;
          s2[j] += dpsi;
          s3[j] += d2psi;
        }
      }
    } else {
      for (int i = tid; i < n_src; i += kT) {
        double u;
        if (!std_geno(i, &u)) continue;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const PairTerms t = pair_terms(ordinal[j], pc_sh[j], u, mp.zz + ((int64_t)i * J + j) * 5, low_sh[j] != 0);
          s1[j] += t.psi;
          s2[j] += t.dpsi;
          s3[j] = fma(t.psi, t.psi, s3[j]);
        }
      }
    }
    double worst = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      double a, b, c;
      if (from_sums) { a = sums_sh[j][0]; b = sums_sh[j][1]; c = sums_sh[j][2]; }
      else { a = block_sum(s1[j], red, tid); b = block_sum(s2[j], red, tid); c = block_sum(s3[j], red, tid); }
      double step;
      if (iter == 0) {  // root of the quadratic model a + b rho + c rho^2 / 2 nearest to the Newton point
        step = b < 0.0 ? -a / b : 0.0;
        const double r1 = step, h = b < 0.0 ? c / (2.0 * b) : 0.0;
        if (fabs(r1 * h) < 0.25) { step = r1 - h * r1 * r1; step = r1 - h * step * step; step = r1 - h * step * step; }
      } else {
        step = b < 0.0 ? -a / b : (c > 0.0 ? a / c : 0.0);
      }
      step = fmax(-0.5, fmin(0.5, step));
      const double next = fmax(-0.99, fmin(0.99, rh[j] + step));
      worst = fmax(worst, fabs(next - rh[j]));
      if (tid == 0) rho_sh[j] = next;
    }
    if (tid == 0 && worst < (iter == 0 ? kSeriesMax : kNewtonStop)) done_sh = 1;
    __syncthreads();
    if (done_sh) break;
  }
  double rh[J];
#pragma unroll
  for (int j = 0; j < J; ++j) rh[j] = rho_sh[j];
  if (tid < J) {
    const double r = rho_sh[tid], is = rsqrt(1.0 - r * r);
    pc_sh[tid][0] = r; pc_sh[tid][1] = is; pc_sh[tid][2] = is * is * is; pc_sh[tid][3] = 3.0 * r * is * is;
    const double dmax = mp.zmax[tid] * (is - 1.0) + fabs(r) * umax * is;
    low_sh[tid] = mp.n_cat[tid] > 0 && dmax <= kLowD && mp.zmax[tid] * dmax + 0.5 * dmax * dmax <= kLowX;
  }
  __syncthreads();
  bool low[J];
#pragma unroll
  for (int j = 0; j < J; ++j) low[j] = low_sh[j] != 0;

  // ---- Gamma pieces.  Phase 1 (thread = person): the SNP-dependent score columns D = (mean, var,
  // psi_j); the A21 rows of the new correlations, sum_i psi_j * d loglik2_i / d (stage-1 parameter
  // of item j), accumulate in a row of shared memory private to the thread (a person touches two
  // threshold entries and the K slopes of each item).  Phase 2 (lane = score column): D' * SC0.
  constexpr int DP = (D + 1) & ~1;                 // Dt row stride (16-byte rows for 128-bit loads)
  constexpr int Q0P = 32 * HALVES;                 // SC0 row stride
  const int vs = mp.n1 | 1;                        // accumulator row stride, odd: conflict-free
  double* Dt = dyn;                                // [kT][DP]
  double* Vt = dyn + kT * DP;                      // [kT][vs]
  double bdd[D * (D + 1) / 2], a21snp[2 * J], ns1[J], ns2[J];  // sums of psi and d psi / d rho, for the closing Newton step
#pragma unroll
  for (int j = 0; j < J; ++j) ns1[j] = ns2[j] = 0.0;
#pragma unroll
  for (int k = 0; k < D * (D + 1) / 2; ++k) bdd[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 2 * J; ++k) a21snp[k] = 0.0;
  double acc[HALVES][D];  // per lane: columns lane (, lane + 32)
#pragma unroll
  for (int h = 0; h < HALVES; ++h)
#pragma unroll
    for (int a = 0; a < D; ++a) acc[h][a] = 0.0;
  const int K = mp.K;
  const double isd = 1.0 / sd, i2v = 1.0 / (2.0 * var);
  double* vrow = Vt + tid * vs;
  for (int b2 = 0; b2 < mp.n1; ++b2) vrow[b2] = 0.0;
  for (int t0 = 0; t0 < n_src; t0 += kT) {
    const int i = t0 + tid;
    double u;
    const bool live = i < n_src && std_geno(i, &u);
    if (live) {
      double d[D], ps[J];
      d[0] = u * isd;
      d[1] = (u * u - 1.0) * i2v;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const double* zz = mp.zz + ((int64_t)i * J + j) * 5;
        const PairTerms t = pair_terms(ordinal[j], pc_sh[j], u, zz, low[j]);
        d[2 + j] = t.psi;
        ns1[j] += t.psi;
        ns2[j] += t.dpsi;
        a21snp[2 * j + 0] -= t.psi * t.dl_du * isd;
        a21snp[2 * j + 1] -= t.psi * fma(t.dl_du, u, 1.0) * i2v;
        const double pa = t.psi * t.dl_a, pb = t.psi * t.dl_b;
        double* vj = vrow + mp.col_of[j];
        if (ordinal[j]) {
          const int nth = mp.n_cat[j] - 1, cat = mp.cat[(int64_t)i * J + j];
          if (cat < nth) vj[cat] += pa;
          if (cat > 0) vj[cat - 1] += pb;
          ps[j] = -(pa + pb);
        } else {  // continuous item: intercept, slopes, variance; zz[0] is the standardised residual
          const double isdj = 1.0 / mp.item_sd[j];
          ps[j] = -pa * isdj;
          vj[0] += ps[j];
          vj[1 + K] -= (t.psi + pa * zz[0]) * (0.5 * isdj * isdj);
        }
      }
      const double* xi = mp.X + (int64_t)i * mp.Kp;
      for (int k = 0; k < K; ++k) {
        const double x = xi[k];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          double* sl = vrow + mp.col_of[j] + (ordinal[j] ? mp.n_cat[j] - 1 : 1) + k;
          *sl = fma(x, ps[j], *sl);
        }
      }
      int k = 0;
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = a; b < D; ++b) { bdd[k] = fma(d[a], d[b], bdd[k]); ++k; }
#pragma unroll
      for (int a = 0; a < D; ++a) Dt[tid * DP + a] = d[a];
    } else {
#pragma unroll
      for (int a = 0; a < D; ++a) Dt[tid * DP + a] = 0.0;
    }
    if (DP > D) Dt[tid * DP + D] = 0.0;
    __syncthreads();
    // phase 2: warp w takes people w*32 .. w*32+31 of the tile, four at a time so that the global
    // loads of four score rows are in flight together (SC0 is padded with zero rows to whole tiles;
    // rows of missing and excluded people have D = 0)
    const double* scw = mp.sc0 + (int64_t)(t0 + warp * 32) * Q0P + lane;
    const double* dw = Dt + warp * 32 * DP;
#pragma unroll 2
    for (int pp0 = 0; pp0 < 32; pp0 += 4) {
      double sc[4][HALVES];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int h = 0; h < HALVES; ++h) sc[k][h] = scw[(pp0 + k) * Q0P + 32 * h];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double dpr[DP];
#pragma unroll
        for (int a = 0; a < DP; a += 2) {
          const double2 two = *reinterpret_cast<const double2*>(dw + (pp0 + k) * DP + a);
          dpr[a] = two.x;
          dpr[a + 1] = two.y;
        }
#pragma unroll
        for (int h = 0; h < HALVES; ++h)
#pragma unroll
          for (int a = 0; a < D; ++a) acc[h][a] = fma(dpr[a], sc[k][h], acc[h][a]);
      }
    }
    __syncthreads();
  }
  // ---- reductions
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double a = block_sum(a21snp[2 * j], red, tid), b = block_sum(a21snp[2 * j + 1], red, tid);
    const double score = block_sum(ns1[j], red, tid), curv = block_sum(ns2[j], red, tid);
    const double step = curv < 0.0 ? fmax(-kFinalClamp, fmin(kFinalClamp, -score / curv)) : 0.0;
    if (tid == 0) { out[mp.o_a21snp + 2 * j] = a; out[mp.o_a21snp + 2 * j + 1] = b; out[mp.o_rho + j] = rh[j] + step; }
  }
  {
    int k = 0;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = a; b < D; ++b) {
        const double v = block_sum(bdd[k++], red, tid);
        if (tid == 0) {
          out[mp.o_bdd + a * D + b] = v;
          out[mp.o_bdd + b * D + a] = v;
          if (a == b && a >= 2) out[mp.o_a22 + (a - 2)] = v;
        }
      }
  }
  // A21 rows: column sums of the threads' private rows
  __syncthreads();
  for (int b2 = warp; b2 < mp.n1; b2 += kT / 32) {
    double v = 0.0;
    for (int r = lane; r < kT; r += 32) v += Vt[r * vs + b2];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) out[mp.o_a21item + b2] = v;
  }
  // lane-owned columns: sum the 8 warps through shared memory
  double* xw = dyn;  // [8][64][D]
  __syncthreads();
#pragma unroll
  for (int h = 0; h < HALVES; ++h)
#pragma unroll
    for (int a = 0; a < D; ++a) xw[((warp * 64) + lane + 32 * h) * D + a] = acc[h][a];
  __syncthreads();
  for (int e = tid; e < 32 * HALVES * D; e += kT) {
    const int b2 = e / D, a = e % D;
    if (b2 >= mp.q0) continue;
    double v = 0.0;
    for (int w = 0; w < kT / 32; ++w) v += xw[(w * 64 + b2) * D + a];
    out[mp.o_bd0 + a * mp.q0 + b2] = v;
  }
}

}  // namespace


// One item count per instantiation; the translation units marg_stats_j*.cu instantiate two each so
// that the ten of them compile in parallel.
template <int JJ>
void launch_marg_stats_j(gwsem_engine* e, const MargProgram& mp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                         int n_src, const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno,
                         int n_pad, double* d_summary, size_t smem) {
#define GW_MARG_L(DOS, HV)                                                                                          \
  {                                                                                                                 \
    GW_CUDA(cudaFuncSetAttribute(marg_stats_kernel<JJ, DOS, HV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    marg_stats_kernel<JJ, DOS, HV><<<n_snps, kT, smem, e->stream>>>(mp, d_rows, pitch, n_src, d_inc8, d_counts, plink1, \
                                                                     d_geno, n_pad, d_summary);                     \
  }
  if (d_geno) {
    if (mp.q0p > 32) GW_MARG_L(true, 2) else GW_MARG_L(true, 1)
  } else {
    if (mp.q0p > 32) GW_MARG_L(false, 2) else GW_MARG_L(false, 1)
  }
#undef GW_MARG_L
}

}  // namespace gwsem
