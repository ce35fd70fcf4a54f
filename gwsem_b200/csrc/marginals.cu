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

struct PairTerms {
  double psi;      // d loglik_i / d rho
  double dl_du;    // d loglik2_i / d u (the standardised genotype)
  double dl_a;     // ordinal item: d/dz1 ; continuous item: d/dv
  double dl_b;     // ordinal item: d/dz0
};

__device__ __forceinline__ PairTerms pair_terms(bool ordinal, double rho, double u, double z1, double z0) {
  PairTerms t;
  if (ordinal) {
    const double s2 = 1.0 - rho * rho, s = sqrt(s2);
    const double a1 = (z1 - rho * u) / s, a0 = (z0 - rho * u) / s;
    const double f1 = dphi(a1), f0 = dphi(a0);   // the +-12 sentinels make the open ends vanish
    const double pi = fmax(rect_prob(a1, a0), 1e-300);
    t.psi = (f1 * (rho * z1 - u) - f0 * (rho * z0 - u)) / (s2 * s * pi);
    t.dl_a = f1 / (s * pi);
    t.dl_b = -f0 / (s * pi);
    t.dl_du = -u - (rho / s) * (f1 - f0) / pi;
  } else {
    const double v = z1, s2 = 1.0 - rho * rho;
    t.psi = (rho * s2 + u * v * (1.0 + rho * rho) - rho * (u * u + v * v)) / (s2 * s2);
    t.dl_du = -(u - rho * v) / s2;
    t.dl_a = -(v - rho * u) / s2;
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
template <int J, bool DOSAGE>
__global__ void __launch_bounds__(kT)
    marg_stats_kernel(MargProgram mp, const uint8_t* __restrict__ rows, int64_t pitch, int n_src,
                      const uint8_t* __restrict__ inc8, const int32_t* __restrict__ counts, int plink1,
                      const double* __restrict__ geno, int n_pad, double* __restrict__ summary) {
  constexpr int D = 2 + J;
  extern __shared__ double dyn[];
  __shared__ double red[kT / 32];
  __shared__ double rho_sh[J];
  __shared__ int done_sh;
  const int snp = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* out = summary + (int64_t)snp * mp.sum_stride;
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

  // ---- people of the model whose call is missing for this SNP (naAction='omit' drops them from
  // every statistic): their item-side score rows, summed and as outer products, let marg_fit.cu
  // down-date the SNP-independent estimates and information to the retained subset.
  if (n < mp.n_incl) {
    double* Mm = dyn;                       // [q0p][q0p], row b owned by lane b / b - 32
    double* msum = dyn + mp.q0p * mp.q0p;   // [q0p]
    for (int k = tid; k < mp.q0p * mp.q0p + mp.q0p; k += kT) dyn[k] = 0.0;
    __syncthreads();
    if (warp == 0) {
      const int halves = mp.q0p / 32;
      for (int i0 = 0; i0 < n_src; i0 += 32) {
        const int i = i0 + lane;
        bool miss = false;
        if (i < n_src && inc8[i]) {
          double u;
          miss = !std_geno(i, &u);
        }
        unsigned todo = __ballot_sync(0xffffffffu, miss);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int64_t ip = i0 + src;
          const double lo = mp.sc0[ip * mp.q0p + lane];
          const double hi = halves > 1 ? mp.sc0[ip * mp.q0p + 32 + lane] : 0.0;
          msum[lane] += lo;
          if (halves > 1) msum[32 + lane] += hi;
          for (int b2 = 0; b2 < mp.q0; ++b2) {
            const double other = b2 < 32 ? __shfl_sync(0xffffffffu, lo, b2) : __shfl_sync(0xffffffffu, hi, b2 - 32);
            Mm[lane * mp.q0p + b2] = fma(lo, other, Mm[lane * mp.q0p + b2]);
            if (halves > 1) Mm[(32 + lane) * mp.q0p + b2] = fma(hi, other, Mm[(32 + lane) * mp.q0p + b2]);
          }
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < mp.q0 * mp.q0; k += kT) out[mp.o_mmat + k] = Mm[(k / mp.q0) * mp.q0p + (k % mp.q0)];
    for (int k = tid; k < mp.q0; k += kT) out[mp.o_msum + k] = msum[k];
    __syncthreads();
  }

  // ---- stage 2: rho_j <- rho_j + sum(psi) / sum(psi^2) until the score vanishes
  if (tid < J) rho_sh[tid] = 0.0;
  if (tid == 0) done_sh = 0;
  __syncthreads();
  for (int iter = 0; iter < 40; ++iter) {
    double s1[J], s2[J], rh[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { s1[j] = s2[j] = 0.0; rh[j] = rho_sh[j]; }
    for (int i = tid; i < n_src; i += kT) {
      double u;
      if (!std_geno(i, &u)) continue;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const PairTerms t = pair_terms(ordinal[j], rh[j], u, mp.z1[(int64_t)i * J + j], mp.z0[(int64_t)i * J + j]);
        s1[j] += t.psi;
        s2[j] = fma(t.psi, t.psi, s2[j]);
      }
    }
    double worst = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double a = block_sum(s1[j], red, tid), b = block_sum(s2[j], red, tid);
      double step = b > 0.0 ? a / b : 0.0;
      step = fmax(-0.5, fmin(0.5, step));
      const double next = fmax(-0.99, fmin(0.99, rh[j] + step));
      worst = fmax(worst, fabs(next - rh[j]));
      if (tid == 0) rho_sh[j] = next;
    }
    if (tid == 0 && worst < 1e-11) done_sh = 1;
    __syncthreads();
    if (done_sh) break;
  }
  double rh[J];
#pragma unroll
  for (int j = 0; j < J; ++j) rh[j] = rho_sh[j];

  // ---- Gamma pieces.  Phase 1 (thread = person): SNP-dependent score columns D = (mean, var, psi_j)
  // and the chain terms; phase 2 (lane = score column): D' * SC0 and the A21 rows.
  double* Dt = dyn;                    // [kT][D]
  double* DLt = dyn + kT * D;          // [kT][J][2]
  double bdd[D * (D + 1) / 2], a21snp[2 * J];
#pragma unroll
  for (int k = 0; k < D * (D + 1) / 2; ++k) bdd[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 2 * J; ++k) a21snp[k] = 0.0;
  double acc[2][D + 1];  // per lane: columns lane, lane + 32; last slot is the A21 item entry
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int a = 0; a <= D; ++a) acc[h][a] = 0.0;
  // role of this lane's columns: which item's stage-1 block, and which entry of it
  int col_item[2], col_sub[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int b = lane + 32 * h;
    col_item[h] = -1;
    col_sub[h] = 0;
    for (int j = 0; j < J; ++j)
      if (b >= mp.col_of[j] && b < mp.col_of[j + 1]) { col_item[h] = j; col_sub[h] = b - mp.col_of[j]; }
  }
  for (int t0 = 0; t0 < n_src; t0 += kT) {
    const int i = t0 + tid;
    double d[D];
#pragma unroll
    for (int a = 0; a < D; ++a) d[a] = 0.0;
    if (i < n_src) {
      double u;
      if (std_geno(i, &u)) {
        d[0] = u / sd;
        d[1] = (u * u - 1.0) / (2.0 * var);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const PairTerms t = pair_terms(ordinal[j], rh[j], u, mp.z1[(int64_t)i * J + j], mp.z0[(int64_t)i * J + j]);
          d[2 + j] = t.psi;
          DLt[(tid * J + j) * 2 + 0] = t.dl_a;
          DLt[(tid * J + j) * 2 + 1] = t.dl_b;
          a21snp[2 * j + 0] += t.psi * (-t.dl_du / sd);
          a21snp[2 * j + 1] += t.psi * (-1.0 / (2.0 * var) - t.dl_du * u / (2.0 * var));
        }
        int k = 0;
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = a; b < D; ++b) { bdd[k] = fma(d[a], d[b], bdd[k]); ++k; }
      }
    }
#pragma unroll
    for (int a = 0; a < D; ++a) Dt[tid * D + a] = d[a];
    if (i >= n_src || (d[0] == 0.0 && d[1] == 0.0))
#pragma unroll
      for (int j = 0; j < J; ++j) DLt[(tid * J + j) * 2] = DLt[(tid * J + j) * 2 + 1] = 0.0;
    __syncthreads();
    // phase 2: warp w takes people w*32 .. w*32+31 of the tile
    for (int pp = 0; pp < 32; ++pp) {
      const int p = warp * 32 + pp;
      const int ip = t0 + p;
      if (ip >= n_src) break;
      const double* dp = Dt + p * D;
      if (dp[0] == 0.0 && dp[1] == 0.0) continue;   // missing / excluded (warp uniform)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int b = lane + 32 * h;
        if (b >= mp.q0p) continue;
        const double sc = mp.sc0[(int64_t)ip * mp.q0p + b];
#pragma unroll
        for (int a = 0; a < D; ++a) acc[h][a] = fma(dp[a], sc, acc[h][a]);
        const int j = col_item[h];
        if (j >= 0) {  // d loglik2 / d (stage-1 parameter b of item j) for the pair (snp, item j)
          const double dla = DLt[(p * J + j) * 2], dlb = DLt[(p * J + j) * 2 + 1];
          double v;
          const int sub = col_sub[h];
          if (mp.n_cat[j] > 0) {
            const int C = mp.n_cat[j], cat = mp.cat[(int64_t)ip * J + j];
            if (sub < C - 1) v = (cat == sub ? dla : 0.0) + (cat == sub + 1 ? dlb : 0.0);
            else v = -mp.X[(int64_t)ip * mp.Kp + (sub - (C - 1))] * (dla + dlb);
          } else {
            const double sdj = mp.item_sd[j], vj = mp.z1[(int64_t)ip * J + j];
            if (sub == 0) v = -dla / sdj;
            else if (sub <= mp.K) v = -mp.X[(int64_t)ip * mp.Kp + (sub - 1)] * dla / sdj;
            else v = -1.0 / (2.0 * sdj * sdj) - dla * vj / (2.0 * sdj * sdj);
          }
          acc[h][D] = fma(dp[2 + j], v, acc[h][D]);
        }
      }
    }
    __syncthreads();
  }
  // ---- reductions
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double a = block_sum(a21snp[2 * j], red, tid), b = block_sum(a21snp[2 * j + 1], red, tid);
    if (tid == 0) { out[mp.o_a21snp + 2 * j] = a; out[mp.o_a21snp + 2 * j + 1] = b; out[mp.o_rho + j] = rh[j]; }
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
  // lane-owned columns: sum the 8 warps through shared memory
  double* xw = dyn;  // [8][64][D+1]
  __syncthreads();
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int a = 0; a <= D; ++a) xw[((warp * 64) + lane + 32 * h) * (D + 1) + a] = acc[h][a];
  __syncthreads();
  for (int e = tid; e < 64 * (D + 1); e += kT) {
    const int b = e / (D + 1), a = e % (D + 1);
    if (b >= mp.q0) continue;
    double v = 0.0;
    for (int w = 0; w < kT / 32; ++w) v += xw[(w * 64 + b) * (D + 1) + a];
    if (a < D) out[mp.o_bd0 + a * mp.q0 + b] = v;
    else if (b < mp.n1) out[mp.o_a21item + b] = v;
  }
}

}  // namespace

void launch_marg_stats(gwsem_engine* e, const MargProgram& mp, const uint8_t* d_rows, int64_t pitch, int n_snps,
                       int n_src, const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno,
                       int n_pad, double* d_summary) {
  if (n_snps <= 0) return;
  const int J = mp.J, D = 2 + J;
  const size_t smem = sizeof(double) * std::max<size_t>(std::max<size_t>((size_t)kT * (D + 2 * J), (size_t)8 * 64 * (D + 1)),
                                                         (size_t)mp.q0p * mp.q0p + mp.q0p);
  e->prof_begin(kFamClassSums);
#define GW_MARG(JJ)                                                                                                  \
  case JJ:                                                                                                           \
    if (d_geno) {                                                                                                    \
      GW_CUDA(cudaFuncSetAttribute(marg_stats_kernel<JJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      marg_stats_kernel<JJ, true><<<n_snps, kT, smem, e->stream>>>(mp, d_rows, pitch, n_src, d_inc8, d_counts, plink1, \
                                                                    d_geno, n_pad, d_summary);                       \
    } else {                                                                                                         \
      GW_CUDA(cudaFuncSetAttribute(marg_stats_kernel<JJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      marg_stats_kernel<JJ, false><<<n_snps, kT, smem, e->stream>>>(mp, d_rows, pitch, n_src, d_inc8, d_counts, plink1, \
                                                                     d_geno, n_pad, d_summary);                      \
    }                                                                                                                \
    break;
  switch (J) {
    GW_MARG(1) GW_MARG(2) GW_MARG(3) GW_MARG(4) GW_MARG(5) GW_MARG(6) GW_MARG(7) GW_MARG(8)
    default: fail("marginals: %d items (1..%d supported)", J, kMargMaxItems);
  }
#undef GW_MARG
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamClassSums);
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
