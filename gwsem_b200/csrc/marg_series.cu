// Series path of the WLS "marginals" statistics for hardcall SNPs without a missing call.
//
// The standardised genotype u takes three values per SNP, and by Mehler's expansion the rectangle probability of an
// ordinal item given u is  pi(rho, u) = pi(0) * sum_k rho^k He_k(u) C_k / k!  with per-person constants
// C_k = (He_{k-1}(z0) phi(z0) - He_{k-1}(z1) phi(z1)) / pi(0).  Every per-person term of the SNP-dependent rows of
// Muthen's A and B (the polyserial score psi, its products with the item-side score columns, the mixed second
// derivatives) is therefore a power series in rho whose coefficients are polynomials in u with SNP-independent
// per-person coefficients, and a sum over people is a combination of GENOTYPE-CLASS SUMS of those coefficients.  The
// class sums come exactly (fixed point, int8 digits) from the tcgen05 class-sum kernel for the whole batch; this file
// holds (i) the host builder of the per-person features and (ii) the per-SNP assembly kernel, which writes the same
// summary record as marg_stats_kernel.  No pass over people per SNP.
//
// Truncation: weighted sums at order rho^3, the unweighted score sum (whose root is the estimate) at order rho^5.  A SNP
// is accepted when the first neglected term of psi, rho^4 * rms over people of its coefficient, is below
// tol * rms(psi at rho = 0) -- otherwise (causal SNPs, very rare alleles) marg_stats_kernel handles it as before.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

#include "common.h"
#include "kernels.h"
#include "marg_series.h"
#include "marginals.h"

namespace gwsem {

using namespace series;

namespace {

// ------------------------------------------------------------------------------------------------ host: features

constexpr int kM = kOrderRoot;   // orders 0..5 carried on the host
struct Ser { double c[kM + 1][kL]; };

// probabilists' Hermite polynomials He_0 .. He_6, ascending coefficients
const double kHe[7][7] = {{1, 0, 0, 0, 0, 0, 0},      {0, 1, 0, 0, 0, 0, 0},     {-1, 0, 1, 0, 0, 0, 0}, {0, -3, 0, 1, 0, 0, 0},
                          {3, 0, -6, 0, 1, 0, 0},     {0, 15, 0, -10, 0, 1, 0},  {-15, 0, 45, 0, -15, 0, 1}};
const double kFact[7] = {1, 1, 2, 6, 24, 120, 720};
inline double he_val(int k, double x) {
  double v = 0.0;
  for (int i = k; i >= 0; --i) v = v * x + kHe[k][i];
  return v;
}
inline void ser_zero(Ser& s) { std::memset(&s, 0, sizeof(s)); }
// q = a * b truncated at order M
inline void ser_mul(const Ser& a, const Ser& b, int M, Ser& q) {
  ser_zero(q);
  for (int m1 = 0; m1 <= M; ++m1)
    for (int l1 = 0; l1 < kL; ++l1) {
      const double x = a.c[m1][l1];
      if (x == 0.0) continue;
      for (int m2 = 0; m1 + m2 <= M; ++m2)
        for (int l2 = 0; l1 + l2 < kL; ++l2) q.c[m1 + m2][l1 + l2] += x * b.c[m2][l2];
    }
}
// q = a / r for a series r whose constant term is exactly 1
inline void ser_div(const Ser& a, const Ser& r, int M, Ser& q) {
  ser_zero(q);
  for (int m = 0; m <= M; ++m) {
    double acc[kL];
    for (int l = 0; l < kL; ++l) acc[l] = a.c[m][l];
    for (int k = 1; k <= m; ++k)
      for (int l1 = 0; l1 < kL; ++l1) {
        const double x = r.c[k][l1];
        if (x == 0.0) continue;
        for (int l2 = 0; l1 + l2 < kL; ++l2) acc[l1 + l2] -= x * q.c[m - k][l2];
      }
    for (int l = 0; l < kL; ++l) q.c[m][l] = acc[l];
  }
}

constexpr int kPer = kPsiAll + kF3 + 2 * kPsiB + 3;   // doubles kept per person and item: psi, F3, F4, F5, rho^4 terms of psi
constexpr int kOffF3 = kPsiAll, kOffF4 = kPsiAll + kF3, kOffF5 = kOffF4 + kPsiB, kOffE = kOffF5 + kPsiB;

// the series of one person and ordinal item from zz = {z1, z0, phi(z1), phi(z0), Phi(z1) - Phi(z0)}
void person_item_series(const double* zz, double* out) {
  const double f1 = zz[2], f0 = zz[3], ip = 1.0 / zz[4];
  const double z1 = f1 == 0.0 ? 0.0 : zz[0], z0 = f0 == 0.0 ? 0.0 : zz[1];   // open categories: the bound carries no density
  double C[7];
  C[0] = 1.0;
  for (int k = 1; k <= 6; ++k) C[k] = (he_val(k - 1, z0) * f0 - he_val(k - 1, z1) * f1) * ip;
  Ser R, Rr, U, Z1, Z0, psi, t, F3, F4, F5;
  ser_zero(R); ser_zero(Rr); ser_zero(U); ser_zero(Z1); ser_zero(Z0);
  for (int k = 0; k <= kM; ++k)
    for (int l = 0; l <= k; ++l) {
      R.c[k][l] = kHe[k][l] * C[k] / kFact[k];
      Rr.c[k][l] = 0.0;
    }
  for (int m = 0; m <= kM; ++m)
    for (int l = 0; l <= m + 1; ++l) Rr.c[m][l] = kHe[m + 1][l] * C[m + 1] / kFact[m];
  for (int k = 1; k <= kOrderB; ++k)
    for (int l = 0; l <= k - 1; ++l) U.c[k][l] = kHe[k - 1][l] * C[k] / kFact[k - 1];
  for (int k = 0; k <= kOrderB; ++k) {
    const double a1 = f1 * ip * he_val(k, z1) / kFact[k], a0 = -f0 * ip * he_val(k, z0) / kFact[k];
    for (int l = 0; l <= k; ++l) { Z1.c[k][l] = kHe[k][l] * a1; Z0.c[k][l] = kHe[k][l] * a0; }
  }
  ser_div(Rr, R, kM, psi);
  ser_div(U, R, kOrderB, t);  ser_mul(psi, t, kOrderB, F3);
  ser_div(Z1, R, kOrderB, t); ser_mul(psi, t, kOrderB, F4);
  ser_div(Z0, R, kOrderB, t); ser_mul(psi, t, kOrderB, F5);
  for (int p = 0; p < kPsiAll; ++p) out[p] = psi.c[psi_m(p)][psi_l(p)];
  for (int p = 0; p < kF3; ++p) out[kOffF3 + p] = F3.c[f3_m(p)][f3_l(p)];
  for (int p = 0; p < kPsiB; ++p) {
    out[kOffF4 + p] = F4.c[psi_m(p)][psi_l(p)];
    out[kOffF5 + p] = F5.c[psi_m(p)][psi_l(p)];
  }
  out[kOffE + 0] = psi.c[4][5]; out[kOffE + 1] = psi.c[4][3]; out[kOffE + 2] = psi.c[4][1];
}

inline int psi_index(int m, int l) {
  for (int p = 0; p < kPsiB; ++p)
    if (psi_m(p) == m && psi_l(p) == l) return p;
  return -1;
}

struct Feature { int type, j, a, p; };   // 0 psi * weight a (0: 1, 1 + b: SC0 column b); 1 psi orders 4, 5; 2 F3; 3 threshold a;
                                         // 4 covariate a; 5 psi_j^2; 6 psi_j psi_a (j > a); 7 SC0 column a

template <typename Body>
void parallel_for(int n, Body body) {
  const int threads = (int)std::min<unsigned>(std::max(1, n), std::max(1u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err(threads);
  auto work = [&](int t) {
    try {
      for (int c = t; c < n; c += threads) body(c);
    } catch (...) { err[t] = std::current_exception(); }
  };
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
}

}  // namespace


// Builds the per-person features in genotype-file order and packs them as int8 digits for the tcgen05 class-sum kernel.
// zz [n_pad][J][5], cat [n_pad][J], X [n_pad][Kp], sc0 [.][q0p] as laid out for marg_stats_kernel; inc[s] = person enters
// the model.  Fills the se_* offsets of mp; returns the packed digits, the grid of every feature and its total.
void build_series_features(MargProgram& mp, int n_src, int n_pad, const std::vector<uint8_t>& inc,
                           const std::vector<double>& zz, const std::vector<uint8_t>& cat, const std::vector<double>& X,
                           const std::vector<double>& sc0, std::vector<int8_t>& packed, std::vector<double>& inv_scale,
                           std::vector<double>& tot, std::vector<double>& err) {
  const int J = mp.J, K = mp.K, q0 = mp.q0;
  // The score polynomial (psi with weight 1, all fifteen terms: its root is the estimate) keeps four base-256 digits of a
  // 2^-30 grid and goes into wide chunks; every other feature feeds the SNP-dependent rows of Muthen's A and B, which
  // need about 1e-6 of their scale: three digits of a 2^-22 grid (rounding error of a class sum over 50k people:
  // < 1e-5 of the sum's own sampling scale even for heavy-tailed features), narrow chunks.
  std::vector<Feature> feats;
  for (int j = 0; j < J; ++j) {
    mp.se_o_psi0[j] = (int)feats.size();
    for (int p = 0; p < kPsiB; ++p) feats.push_back({0, j, 0, p});
    mp.se_o_hi[j] = (int)feats.size();
    for (int p = kPsiB; p < kPsiAll; ++p) feats.push_back({1, j, 0, p});
  }
  while (feats.size() % kTcWideFeatures) feats.push_back({-1, 0, 0, 0});
  const int n_wide = (int)feats.size();
  for (int j = 0; j < J; ++j) {
    mp.se_o_psi[j] = (int)feats.size();
    for (int w = 1; w <= q0; ++w)
      for (int p = 0; p < kPsiB; ++p) feats.push_back({0, j, w, p});
    mp.se_o_f3[j] = (int)feats.size();
    for (int p = 0; p < kF3; ++p) feats.push_back({2, j, 0, p});
    mp.se_o_thr[j] = (int)feats.size();
    for (int t = 0; t + 1 < mp.n_cat[j]; ++t)
      for (int p = 0; p < kPsiB; ++p) feats.push_back({3, j, t, p});
    mp.se_o_slope[j] = (int)feats.size();
    for (int k = 0; k < K; ++k)
      for (int p = 0; p < kPsiB; ++p) feats.push_back({4, j, k, p});
    mp.se_o_diag[j] = (int)feats.size();
    for (int p = 0; p < kDiag; ++p) feats.push_back({5, j, 0, p});
  }
  mp.se_o_cross = (int)feats.size();
  for (int b = 0; b < J; ++b)
    for (int a = b + 1; a < J; ++a)
      for (int p = 0; p < kCross; ++p) feats.push_back({6, a, b, p});
  mp.se_o_sc0 = (int)feats.size();
  for (int b = 0; b < q0; ++b) feats.push_back({7, 0, b, 0});
  while ((feats.size() - n_wide) % kTcNarrowFeatures) feats.push_back({-1, 0, 0, 0});
  const int nf = (int)feats.size(), n_narrow = nf - n_wide;
  const int wide_chunks = n_wide / kTcWideFeatures, chunks = wide_chunks + n_narrow / kTcNarrowFeatures;
  mp.se_nf = nf; mp.se_n_wide = n_wide; mp.se_n_narrow = n_narrow;

  // per person and item: the series
  std::vector<double> ps((size_t)n_pad * J * kPer, 0.0);
  parallel_for(64, [&](int c) {
    const int lo = (int)((int64_t)n_src * c / 64), hi = (int)((int64_t)n_src * (c + 1) / 64);
    for (int s = lo; s < hi; ++s) {
      if (!inc[s]) continue;
      for (int j = 0; j < J; ++j) person_item_series(&zz[((size_t)s * J + j) * 5], &ps[((size_t)s * J + j) * kPer]);
    }
  });
  // acceptance bound: second moments of the rho^4 coefficients of psi (u^5, u^3, u^1) and of C1
  err.assign((size_t)J * 10, 0.0);
  {
    double cnt = 0.0;
    for (int s = 0; s < n_src; ++s) cnt += inc[s] ? 1.0 : 0.0;
    for (int j = 0; j < J; ++j) {
      long double m[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int s = 0; s < n_src; ++s) {
        if (!inc[s]) continue;
        const double* e = &ps[((size_t)s * J + j) * kPer];
        const double q5 = e[kOffE], q3 = e[kOffE + 1], q1 = e[kOffE + 2], c1 = e[0];
        m[0] += q5 * q5; m[1] += q5 * q3; m[2] += q5 * q1; m[3] += q3 * q3; m[4] += q3 * q1; m[5] += q1 * q1; m[6] += c1 * c1;
      }
      for (int k = 0; k < 7; ++k) err[(size_t)j * 10 + k] = (double)(m[k] / cnt);
    }
  }

  const size_t wide_bytes = tc5_wide_chunk_bytes(n_pad), narrow_bytes = tc5_narrow_chunk_bytes(n_pad);
  packed.assign(wide_bytes * wide_chunks + narrow_bytes * (chunks - wide_chunks), 0);
  inv_scale.assign(nf, 0.0);
  tot.assign(nf, 0.0);
  const int q0p = mp.q0p, Kp = mp.Kp;
  std::atomic<bool> heavy_tail(false);
  int n_incl_people = 0;
  for (int s = 0; s < n_src; ++s) n_incl_people += inc[s] ? 1 : 0;
  constexpr int kParts = 4;   // a chunk's features in four tasks: 92 tasks keep 16 host threads evenly busy
  parallel_for(chunks * kParts, [&](int task) {
    const int c = task / kParts, part = task % kParts;
    const bool narrow = c >= wide_chunks;
    const int cf_all = narrow ? kTcNarrowFeatures : kTcWideFeatures, cf = cf_all / kParts, fi_base = part * cf;
    const int first = (narrow ? n_wide + (c - wide_chunks) * kTcNarrowFeatures : c * kTcWideFeatures) + fi_base;
    const int grid_bits = narrow ? 22 : 30;
    int8_t* dst = packed.data() + (narrow ? wide_bytes * wide_chunks + narrow_bytes * (size_t)(c - wide_chunks) : wide_bytes * (size_t)c);
    std::vector<double> h((size_t)cf * n_pad, 0.0);
    std::vector<int64_t> q((size_t)cf * n_pad, 0);
    for (int fi = 0; fi < cf; ++fi) {
      const Feature f = feats[first + fi];
      if (f.type < 0) continue;
      double* hv = &h[(size_t)fi * n_pad];
      CrossTerm ct = {0, 0, 0};
      if (f.type == 6) ct = cross_term(f.p);
      for (int s = 0; s < n_src; ++s) {
        if (!inc[s]) continue;
        const double* e = &ps[((size_t)s * J + f.j) * kPer];
        double v = 0.0;
        switch (f.type) {
          case 0: v = e[f.p] * (f.a == 0 ? 1.0 : sc0[(size_t)s * q0p + f.a - 1]); break;
          case 1: v = e[f.p]; break;
          case 2: v = e[kOffF3 + f.p]; break;
          case 3: {
            const int ct2 = cat[(size_t)s * J + f.j];
            v = (ct2 == f.a ? e[kOffF4 + f.p] : 0.0) + (ct2 == f.a + 1 ? e[kOffF5 + f.p] : 0.0);
            break;
          }
          case 4: v = (e[kOffF4 + f.p] + e[kOffF5 + f.p]) * X[(size_t)s * Kp + f.a]; break;
          case 5: {
            const int m = diag_m(f.p), l = diag_l(f.p);
            for (int m1 = 0; m1 <= m; ++m1)
              for (int l1 = 0; l1 <= l; ++l1) {
                const int p1 = psi_index(m1, l1), p2 = psi_index(m - m1, l - l1);
                if (p1 >= 0 && p2 >= 0) v += e[p1] * e[p2];
              }
            break;
          }
          case 6: {
            const double* eb = &ps[((size_t)s * J + f.a) * kPer];
            for (int l1 = 0; l1 <= ct.l; ++l1) {
              const int p1 = psi_index(ct.ma, l1), p2 = psi_index(ct.mb, ct.l - l1);
              if (p1 >= 0 && p2 >= 0) v += e[p1] * eb[p2];
            }
            break;
          }
          case 7: v = sc0[(size_t)s * q0p + f.a]; break;
        }
        hv[s] = v;
      }
      double mx = 0.0, ss = 0.0;
      for (int s = 0; s < n_src; ++s) { mx = std::max(mx, std::fabs(hv[s])); ss += hv[s] * hv[s]; }
      // three digits resolve 2^-23 of the largest value: for a feature whose largest value is more than a thousand
      // times its root mean square (an outlier in a covariate, say) that is no longer 1e-5 of a class sum's scale
      if (narrow && ss > 0.0 && mx > 1000.0 * std::sqrt(ss / std::max(1, n_incl_people))) heavy_tail.store(true);
      int ex = 0;
      if (mx > 0) std::frexp(mx, &ex);
      const double sc = std::ldexp(1.0, grid_bits - ex);
      inv_scale[first + fi] = std::ldexp(1.0, ex - grid_bits);
      long double acc = 0;
      int64_t* qv = &q[(size_t)fi * n_pad];
      for (int s = 0; s < n_src; ++s) {
        qv[s] = std::llrint(hv[s] * sc);
        acc += (long double)qv[s];
      }
      tot[first + fi] = (double)acc * inv_scale[first + fi];
    }
    tc5_pack_chunk(q.data(), fi_base, cf, n_pad, dst, narrow);
  });
  if (heavy_tail.load()) {   // leave such models to the per-person kernel
    mp.se_nf = mp.se_n_wide = mp.se_n_narrow = 0;
    packed.clear();
  }
}

// ------------------------------------------------------------------------------------------------ device: assembly

namespace {

constexpr int kWarps = 4;

struct WarpShared {
  double rp[kMargMaxItems][kM + 1];   // rho_j^m
  double u0[kL], ud1[kL], ud2[kL];    // u_0^l, u_1^l - u_0^l, u_2^l - u_0^l
  double psum[kMargMaxItems][3];      // sum of psi_j over the people of each class
  double fsum[kMargMaxItems][3];      // sum of psi_j * d log pi / du
};

// sum over all people of a feature group: terms (m(p), l(p)), features f0 .. f0 + n - 1 (class 0 = total - het - hom)
template <typename MF, typename LF>
__device__ __forceinline__ double eval_terms(const WarpShared& w, const double* rp, const double* __restrict__ T,
                                             const double* __restrict__ S1, const double* __restrict__ S2, int f0, int n,
                                             MF mf, LF lf) {
  double v = 0.0;
  for (int p = 0; p < n; ++p) {
    const int l = lf(p);
    v += rp[mf(p)] * (w.u0[l] * T[f0 + p] + w.ud1[l] * S1[f0 + p] + w.ud2[l] * S2[f0 + p]);
  }
  return v;
}

__global__ void __launch_bounds__(kWarps * 32)
    marg_series_kernel(MargProgram mp, const int32_t* __restrict__ counts, int n_snps, double* __restrict__ summary,
                       double tol2, double rho_max) {
  __shared__ WarpShared sh[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int snp = blockIdx.x * kWarps + warp;
  if (snp >= n_snps) return;
  WarpShared& w = sh[warp];
  const int J = mp.J, D = 2 + J, q0 = mp.q0, K = mp.K, nf = mp.se_nf;
  const int32_t* c = counts + 8 * (int64_t)snp;
  const int n = mp.n_incl - c[2];
  bool ok = n == mp.n_incl && n >= 2;
  const double dn = (double)n, mu = ((double)c[0] + 2.0 * (double)c[1]) / dn;
  const double ncls[3] = {dn - (double)c[0] - (double)c[1], (double)c[0], (double)c[1]};
  const double var = (ncls[0] * mu * mu + ncls[1] * (1.0 - mu) * (1.0 - mu) + ncls[2] * (2.0 - mu) * (2.0 - mu)) / dn;
  ok = ok && var > 0.0;
  if (!ok) {
    if (lane == 0) mp.se_done[snp] = 0;
    return;
  }
  const double sd = sqrt(var), isd = 1.0 / sd, i2v = 1.0 / (2.0 * var);
  const double uc[3] = {(0.0 - mu) * isd, (1.0 - mu) * isd, (2.0 - mu) * isd};
  const double* __restrict__ S1 = mp.se_sums + ((int64_t)snp * 2 + 0) * nf;
  const double* __restrict__ S2 = mp.se_sums + ((int64_t)snp * 2 + 1) * nf;
  const double* __restrict__ T = mp.se_tot;
  if (lane < kL) {
    double p0 = 1.0, p1 = 1.0, p2 = 1.0;
    for (int l = 0; l < lane; ++l) { p0 *= uc[0]; p1 *= uc[1]; p2 *= uc[2]; }
    w.u0[lane] = p0; w.ud1[lane] = p1 - p0; w.ud2[lane] = p2 - p0;
  }
  __syncwarp();
  // ---- the correlation of every item: root of the score polynomial sum_m rho^m a_m
  bool good = true, start_ok = true;
  if (lane < J) {
    const int j = lane;
    double a[kM + 1];
    for (int m = 0; m <= kM; ++m) a[m] = 0.0;
    for (int p = 0; p < kPsiAll; ++p) {
      const int f = p < kPsiB ? mp.se_o_psi0[j] + p : mp.se_o_hi[j] + (p - kPsiB), l = psi_l(p);
      a[psi_m(p)] += w.u0[l] * T[f] + w.ud1[l] * S1[f] + w.ud2[l] * S2[f];
    }
    double rho = 0.0;
    bool conv = false;
    for (int it = 0; it < 40; ++it) {
      double P = a[kM], dP = 0.0;
      for (int m = kM - 1; m >= 0; --m) { dP = fma(dP, rho, P); P = fma(P, rho, a[m]); }
      if (!(dP < 0.0)) break;
      const double step = -P / dP;
      rho += step;
      if (!(fabs(rho) <= rho_max)) break;
      if (fabs(step) < 1e-15) { conv = true; break; }
    }
    good = conv;
    double rel4 = 0.0;
    if (good) {
      // first neglected term of psi against psi at rho = 0, in mean square over people
      const double* E = mp.se_err + 10 * j;
      double e2 = 0.0;
      for (int cls = 0; cls < 3; ++cls) {
        const double u2 = uc[cls] * uc[cls];
        const double poly = u2 * (E[5] + u2 * (2.0 * E[4] + u2 * (2.0 * E[2] + E[3] + u2 * (2.0 * E[1] + u2 * E[0]))));
        e2 += ncls[cls] / dn * poly;
      }
      const double r2 = rho * rho, r8 = r2 * r2 * r2 * r2;
      good = r8 * e2 <= tol2 * E[6];
      rel4 = sqrt(r8 * e2 / E[6]);
    }
    // The root itself comes from the order-5 series: its first neglected term is about 1.5 rho^2 times the order-4 one
    // (measured on the C3 items, scripts/series_proto.py).  When that is small the root is a Newton starting point
    // within marg_stats_kernel's own stopping distance, and that kernel goes straight to its Gamma pass.
    start_ok = conv && 1.5 * rho * rho * rel4 <= 1e-5;
    double pw = 1.0;
    for (int m = 0; m <= kM; ++m) { w.rp[j][m] = pw; pw *= rho; }
  }
  good = __all_sync(0xffffffffu, good);
  if (!good) {
    start_ok = __all_sync(0xffffffffu, start_ok);
    if (start_ok && lane < J) summary[(int64_t)snp * mp.sum_stride + mp.o_rho + lane] = w.rp[lane][1];
    if (lane == 0) mp.se_done[snp] = start_ok ? 2 : 0;
    return;
  }
  __syncwarp();
  double* out = summary + (int64_t)snp * mp.sum_stride;
  for (int k = lane; k < mp.sum_stride; k += 32) out[k] = 0.0;
  __syncwarp();
  // per class sums of psi_j and of psi_j * d log pi / du (lane = item * 3 + class)
  for (int e = lane; e < 3 * J; e += 32) {
    const int j = e / 3, cls = e % 3;
    double ps = 0.0, fs = 0.0;
    for (int p = 0; p < kPsiAll; ++p) {
      const int f = p < kPsiB ? mp.se_o_psi0[j] + p : mp.se_o_hi[j] + (p - kPsiB);
      const double s = cls == 0 ? T[f] - S1[f] - S2[f] : (cls == 1 ? S1[f] : S2[f]);
      double up = 1.0;
      for (int l = 0; l < psi_l(p); ++l) up *= uc[cls];
      ps += w.rp[j][psi_m(p)] * up * s;
    }
    for (int p = 0; p < kF3; ++p) {
      const int f = mp.se_o_f3[j] + p;
      const double s = cls == 0 ? T[f] - S1[f] - S2[f] : (cls == 1 ? S1[f] : S2[f]);
      double up = 1.0;
      for (int l = 0; l < f3_l(p); ++l) up *= uc[cls];
      fs += w.rp[j][f3_m(p)] * up * s;
    }
    w.psum[j][cls] = ps;
    w.fsum[j][cls] = fs;
  }
  __syncwarp();
  if (lane == 0) {
    out[0] = dn; out[1] = mu; out[2] = var;
    double b00 = 0.0, b01 = 0.0, b11 = 0.0;
    for (int cls = 0; cls < 3; ++cls) {
      const double d0 = uc[cls] * isd, d1 = (uc[cls] * uc[cls] - 1.0) * i2v;
      b00 += ncls[cls] * d0 * d0; b01 += ncls[cls] * d0 * d1; b11 += ncls[cls] * d1 * d1;
    }
    out[mp.o_bdd + 0] = b00; out[mp.o_bdd + 1] = b01; out[mp.o_bdd + D] = b01; out[mp.o_bdd + D + 1] = b11;
  }
  if (lane < J) {
    const int j = lane;
    out[mp.o_rho + j] = w.rp[j][1];
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    for (int cls = 0; cls < 3; ++cls) {
      const double u = uc[cls], t = w.fsum[j][cls] - u * w.psum[j][cls];   // sum of psi * dl_du over the class
      a0 -= t * isd;
      a1 -= (u * t + w.psum[j][cls]) * i2v;
      b0 += u * isd * w.psum[j][cls];
      b1 += (u * u - 1.0) * i2v * w.psum[j][cls];
    }
    out[mp.o_a21snp + 2 * j] = a0; out[mp.o_a21snp + 2 * j + 1] = a1;
    out[mp.o_bdd + 0 * D + 2 + j] = b0; out[mp.o_bdd + (2 + j) * D + 0] = b0;
    out[mp.o_bdd + 1 * D + 2 + j] = b1; out[mp.o_bdd + (2 + j) * D + 1] = b1;
    const double dg = eval_terms(w, w.rp[j], T, S1, S2, mp.se_o_diag[j], kDiag, [](int p) { return diag_m(p); }, [](int p) { return diag_l(p); });
    out[mp.o_bdd + (2 + j) * D + 2 + j] = dg;
    out[mp.o_a22 + j] = dg;
  }
  // psi_a psi_b, a > b
  for (int e = lane; e < J * (J - 1) / 2; e += 32) {
    int b = 0, k = e;
    while (k >= J - 1 - b) { k -= J - 1 - b; ++b; }
    const int a = b + 1 + k, f0 = mp.se_o_cross + e * kCross;
    double v = 0.0;
    for (int p = 0; p < kCross; ++p) {
      const CrossTerm t = cross_term(p);
      v += w.rp[a][t.ma] * w.rp[b][t.mb] * (w.u0[t.l] * T[f0 + p] + w.ud1[t.l] * S1[f0 + p] + w.ud2[t.l] * S2[f0 + p]);
    }
    out[mp.o_bdd + (2 + a) * D + 2 + b] = v;
    out[mp.o_bdd + (2 + b) * D + 2 + a] = v;
  }
  // D' * SC0
  for (int e = lane; e < D * q0; e += 32) {
    const int a = e / q0, b2 = e % q0;
    double v;
    if (a < 2) {
      const int f = mp.se_o_sc0 + b2;
      const double s[3] = {T[f] - S1[f] - S2[f], S1[f], S2[f]};
      v = 0.0;
      for (int cls = 0; cls < 3; ++cls) v += (a == 0 ? uc[cls] * isd : (uc[cls] * uc[cls] - 1.0) * i2v) * s[cls];
    } else {
      const int j = a - 2;
      v = eval_terms(w, w.rp[j], T, S1, S2, mp.se_o_psi[j] + b2 * kPsiB, kPsiB, [](int p) { return psi_m(p); }, [](int p) { return psi_l(p); });
    }
    out[mp.o_bd0 + a * q0 + b2] = v;
  }
  // A21 rows of the new correlations against the items' stage-1 parameters
  for (int e = lane; e < mp.n1; e += 32) {
    int j = 0;
    while (j + 1 < J && e >= mp.col_of[j + 1]) ++j;
    const int r = e - mp.col_of[j], nth = mp.n_cat[j] - 1;
    double v;
    if (r < nth)
      v = eval_terms(w, w.rp[j], T, S1, S2, mp.se_o_thr[j] + r * kPsiB, kPsiB, [](int p) { return psi_m(p); }, [](int p) { return psi_l(p); });
    else
      v = -eval_terms(w, w.rp[j], T, S1, S2, mp.se_o_slope[j] + (r - nth) * kPsiB, kPsiB, [](int p) { return psi_m(p); }, [](int p) { return psi_l(p); });
    out[mp.o_a21item + e] = v;
  }
  (void)K;
  if (lane == 0) mp.se_done[snp] = 1;
}

}  // namespace

void launch_marg_series(gwsem_engine* e, const MargProgram& mp, const int32_t* d_counts, int n_snps, double* d_summary) {
  if (n_snps <= 0) return;
  static const double tol = getenv("GWSEM_SERIES_TOL") ? atof(getenv("GWSEM_SERIES_TOL")) : 3e-6;      // tuning aids
  static const double rho_max = getenv("GWSEM_SERIES_RHOMAX") ? atof(getenv("GWSEM_SERIES_RHOMAX")) : 0.15;
  e->prof_begin(kFamMargSeries);
  marg_series_kernel<<<(n_snps + kWarps - 1) / kWarps, kWarps * 32, 0, e->stream>>>(mp, d_counts, n_snps, d_summary, tol * tol, rho_max);
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamMargSeries);
  e->launches++;
}

}  // namespace gwsem
