// SNP-independent half of the WLS "marginals" summary statistics (reference R/model.R:5,
// mxFitFunctionWLS(allContinuousMethod="marginals")), computed once per
// model on the host: stage 1 for every item given the exogenous covariates (OLS / ordered
// probit), stage 2 for every item pair (Pearson / polyserial / polychoric), and the per-person
// score vectors the per-SNP kernels need to finish Muthen's (1984) asymptotic covariance.
// The SNP-dependent half (snp mean / variance, snp-item correlations, their scores) is
// marginals.cu.  Mirrors oracle/marginals_oracle.py.
#include "marginals.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <thread>
#include <vector>

#include "common.h"

namespace gwsem {

namespace {

// Runs body(chunk, begin, end) over a fixed partition of [0, n) into kChunks ranges on up to
// kChunks host threads.  The partition does not depend on the machine, so sums that are combined
// in chunk order are reproducible from run to run and from host to host.
constexpr int kChunks = 32;
template <typename Body>
void for_chunks(int n, Body body) {
  const int threads = (int)std::min<unsigned>(kChunks, std::max(1u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err(threads);
  auto work = [&](int t) {
    try {
      for (int c = t; c < kChunks; c += threads) body(c, (int)((int64_t)n * c / kChunks), (int)((int64_t)n * (c + 1) / kChunks));
    } catch (...) { err[t] = std::current_exception(); }
  };
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
}

constexpr double kZmax = 12.0;
const double kInvSqrt2Pi = 0.3989422804014327;

inline double phi(double z) { return kInvSqrt2Pi * std::exp(-0.5 * z * z); }
inline double Phi(double z) { return 0.5 * std::erfc(-z * 0.7071067811865476); }

// 32-point Gauss-Legendre on [-1, 1]
const double kGLx[16] = {0.0483076656877383162, 0.1444719615827964934, 0.2392873622521370745, 0.3318686022821276498,
                         0.4213512761306353454, 0.5068999089322293900, 0.5877157572407623290, 0.6630442669302152010,
                         0.7321821187402896804, 0.7944837959679424070, 0.8493676137325699701, 0.8963211557660521240,
                         0.9349060759377396892, 0.9647622555875064308, 0.9856115115452683354, 0.9972638618494815635};
const double kGLw[16] = {0.0965400885147278006, 0.0956387200792748594, 0.0938443990808045654, 0.0911738786957638847,
                         0.0876520930044038111, 0.0833119242269467552, 0.0781938957870703065, 0.0723457941088485062,
                         0.0658222227763618468, 0.0586840934785355471, 0.0509980592623761762, 0.0428358980222266807,
                         0.0342738629130214331, 0.0253920653092620595, 0.0162743947309056706, 0.0070186100094700966};

// P(X <= h, Y <= k) for the standard bivariate normal (Drezner & Wesolowsky 1990 form)
double bvn_cdf(double h, double k, double rho) {
  const double a = std::asin(rho);
  double sum = 0.0;
  for (int i = 0; i < 16; ++i)
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const double t = 0.5 * a * (sgn * kGLx[i] + 1.0);
      const double st = std::sin(t), ct = std::cos(t);
      sum += kGLw[i] * std::exp(-(h * h + k * k - 2 * h * k * st) / (2 * ct * ct));
    }
  return Phi(h) * Phi(k) + 0.5 * a * sum / (2 * M_PI);
}
double bvn_pdf(double h, double k, double rho) {
  const double s2 = 1 - rho * rho;
  return std::exp(-(h * h - 2 * rho * h * k + k * k) / (2 * s2)) / (2 * M_PI * std::sqrt(s2));
}

// solve A x = b for symmetric positive definite A (n x n, row major); A is overwritten
bool chol_solve(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * n + j];
      for (int k = 0; k < j; ++k) v -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= A[i * n + k] * b[k];
    b[i] = v / A[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= A[k * n + i] * b[k];
    b[i] = v / A[i * n + i];
  }
  return true;
}

// 1-D root of an increasing-then-decreasing log-likelihood's score: Newton with numeric slope
template <typename F>
double solve_rho(F score, double rho0) {
  double rho = rho0, g = score(rho);
  for (int it = 0; it < 100; ++it) {
    const double h = 1e-6;
    const double hi = std::min(rho + h, 0.999), lo = std::max(rho - h, -0.999);
    const double gp = (score(hi) - score(lo)) / (hi - lo);
    const double step = gp < 0 ? -g / gp : (g > 0 ? 0.1 : -0.1);
    double next = std::min(0.99, std::max(-0.99, rho + step));
    double gn = score(next);
    for (int k = 0; std::fabs(gn) > std::fabs(g) && k < 30; ++k) {
      next = rho + (next - rho) / 2;
      gn = score(next);
    }
    const double moved = std::fabs(next - rho);
    rho = next;
    g = gn;
    if (moved < 1e-13) break;
  }
  return rho;
}

}  // namespace

// Stage 1 of one item on the people in `rows` (indices into the model's data rows).
void MargItem::fit(const std::vector<int>& rows, const std::vector<const double*>& X) {
  const int n = (int)rows.size(), K = (int)X.size();
  if (n_cat == 0) {  // OLS on [1, X], ML variance
    const int d = K + 1;
    std::vector<double> XtX(d * d, 0.0), Xty(d, 0.0), xi(d);
    for (int r : rows) {
      xi[0] = 1;
      for (int k = 0; k < K; ++k) xi[k + 1] = X[k][r];
      for (int a = 0; a < d; ++a) {
        Xty[a] += xi[a] * values[r];
        for (int b = 0; b < d; ++b) XtX[a * d + b] += xi[a] * xi[b];
      }
    }
    if (!chol_solve(XtX, Xty, d)) fail("exogenous covariates are collinear");
    beta = Xty;
    double ss = 0;
    for (int r : rows) {
      double e = values[r] - beta[0];
      for (int k = 0; k < K; ++k) e -= beta[k + 1] * X[k][r];
      ss += e * e;
    }
    var = ss / n;
    n_par = K + 2;
    // exact information of (beta, var) at the MLE: X'X / var, n / (2 var^2); the cross term vanishes
    hess.assign((size_t)n_par * n_par, 0.0);
    for (int r : rows) {
      xi[0] = 1;
      for (int k = 0; k < K; ++k) xi[k + 1] = X[k][r];
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) hess[(size_t)a * n_par + b] += xi[a] * xi[b] / var;
    }
    hess[(size_t)(n_par - 1) * n_par + (n_par - 1)] = n / (2 * var * var);
    return;
  }
  // ordered probit, Newton-Raphson with the exact Hessian
  const int C = n_cat, d = C - 1 + K;
  std::vector<double> par(d, 0.0);
  {
    std::vector<double> cnt(C, 0.0);
    for (int r : rows) cnt[(int)values[r]] += 1;
    double cum = 0;
    for (int c = 0; c + 1 < C; ++c) {
      cum += cnt[c] / n;
      const double p = std::min(1 - 1e-10, std::max(1e-10, cum));
      // inverse normal cdf by bisection (start values only)
      double lo = -10, hi = 10;
      for (int it = 0; it < 80; ++it) {
        const double mid = 0.5 * (lo + hi);
        (Phi(mid) < p ? lo : hi) = mid;
      }
      par[c] = 0.5 * (lo + hi);
    }
  }
  std::vector<double> g(d), H(d * d), j1(d), j0(d);
  bool converged = false;
  for (int it = 0; it < 100; ++it) {
    std::vector<double> gp((size_t)kChunks * d, 0.0), Hp((size_t)kChunks * d * d, 0.0);
    for_chunks(n, [&](int chunk, int lo_i, int hi_i) {
      double* gc = &gp[(size_t)chunk * d];
      double* Hcn = &Hp[(size_t)chunk * d * d];
      std::vector<double> j1(d), j0(d);
      for (int ri = lo_i; ri < hi_i; ++ri) {
        const int r = rows[ri];
        const int c = (int)values[r];
        double eta = 0;
        for (int k = 0; k < K; ++k) eta += par[C - 1 + k] * X[k][r];
        const double z1 = (c < C - 1 ? par[c] : kZmax) - eta, z0 = (c > 0 ? par[c - 1] : -kZmax) - eta;
        const double p1 = c < C - 1 ? phi(z1) : 0.0, p0 = c > 0 ? phi(z0) : 0.0;
        const double pi = Phi(z1) - Phi(z0);
        std::fill(j1.begin(), j1.end(), 0.0);
        std::fill(j0.begin(), j0.end(), 0.0);
        if (c < C - 1) j1[c] = 1;
        if (c > 0) j0[c - 1] = 1;
        for (int k = 0; k < K; ++k) j1[C - 1 + k] = j0[C - 1 + k] = -X[k][r];
        const double d1 = -z1 * p1, d0 = -z0 * p0;
        for (int a = 0; a < d; ++a) {
          const double ga = (p1 * j1[a] - p0 * j0[a]) / pi;
          gc[a] += ga;
          for (int b = 0; b < d; ++b) {
            const double gb = (p1 * j1[b] - p0 * j0[b]) / pi;
            Hcn[a * d + b] += ga * gb - (d1 * j1[a] * j1[b] - d0 * j0[a] * j0[b]) / pi;
          }
        }
      }
    });
    std::fill(g.begin(), g.end(), 0.0);
    std::fill(H.begin(), H.end(), 0.0);
    for (int chunk = 0; chunk < kChunks; ++chunk) {
      for (int a = 0; a < d; ++a) g[a] += gp[(size_t)chunk * d + a];
      for (int e = 0; e < d * d; ++e) H[e] += Hp[(size_t)chunk * d * d + e];
    }
    if (converged) {  // one more evaluation at the estimate: keep its exact information
      hess = H;
      break;
    }
    std::vector<double> step = g;
    std::vector<double> Hc = H;
    if (!chol_solve(Hc, step, d)) fail("ordered probit regression failed (information matrix not positive definite)");
    double mx = 0;
    for (int a = 0; a < d; ++a) {
      par[a] += step[a];
      mx = std::max(mx, std::fabs(step[a]));
    }
    if (mx < 1e-12) converged = true;
  }
  if (hess.empty()) hess = H;
  th.assign(par.begin(), par.begin() + C - 1);
  beta.assign(par.begin() + C - 1, par.end());
  n_par = d;
}

// Everything SNP-independent, for the people in `rows`.
void MarginalsHost::build(const std::vector<int>& rows) {
  const int n = (int)rows.size(), K = (int)X.size(), J = (int)items.size();
  for (auto& it : items) it.fit(rows, X);
  col_of.assign(J + 1, 0);
  for (int j = 0; j < J; ++j) col_of[j + 1] = col_of[j] + items[j].n_par;
  n1 = col_of[J];
  pairs.clear();
  for (int b = 0; b < J; ++b)
    for (int a = b + 1; a < J; ++a) pairs.emplace_back(a, b);  // column-major lower triangle, as the oracle
  q0 = n1 + (int)pairs.size();
  n_rows_total = items.empty() ? 0 : items[0].n_values;
  sc0.assign((size_t)n_rows_total * q0, 0.0);
  z1.assign((size_t)n_rows_total * J, 0.0);
  z0.assign((size_t)n_rows_total * J, 0.0);
  rho.assign(pairs.size(), 0.0);
  A21.assign(pairs.size() * (size_t)n1, 0.0);
  // stage-1 scores and the standardised quantities each person contributes
  for (int j = 0; j < J; ++j) {
    MargItem& it = items[j];
    const int base = col_of[j];
    for (int r : rows) {
      double* sc = &sc0[(size_t)r * q0 + base];
      if (it.n_cat == 0) {
        double e = it.values[r] - it.beta[0];
        for (int k = 0; k < K; ++k) e -= it.beta[k + 1] * X[k][r];
        sc[0] = e / it.var;
        for (int k = 0; k < K; ++k) sc[1 + k] = X[k][r] * e / it.var;
        sc[K + 1] = (e * e - it.var) / (2 * it.var * it.var);
        z1[(size_t)r * J + j] = e / std::sqrt(it.var);  // continuous: standardised residual
      } else {
        const int C = it.n_cat, c = (int)it.values[r];
        double eta = 0;
        for (int k = 0; k < K; ++k) eta += it.beta[k] * X[k][r];
        const double a1 = (c < C - 1 ? it.th[c] : kZmax) - eta, a0 = (c > 0 ? it.th[c - 1] : -kZmax) - eta;
        const double p1 = c < C - 1 ? phi(a1) : 0.0, p0 = c > 0 ? phi(a0) : 0.0, pi = Phi(a1) - Phi(a0);
        if (c < C - 1) sc[c] = p1 / pi;
        if (c > 0) sc[c - 1] -= p0 / pi;
        for (int k = 0; k < K; ++k) sc[C - 1 + k] = X[k][r] * (p0 - p1) / pi;
        z1[(size_t)r * J + j] = a1;
        z0[(size_t)r * J + j] = a0;
      }
    }
  }
  // stage 2 for item pairs
  std::vector<double> da, db;
  for (size_t pk = 0; pk < pairs.size(); ++pk) {
    const int ia = pairs[pk].first, ib = pairs[pk].second;
    MargItem &a = items[ia], &b = items[ib];
    const bool oa = a.n_cat > 0, ob = b.n_cat > 0;
    auto Z1 = [&](int j, int r) { return z1[(size_t)r * J + j]; };
    auto Z0 = [&](int j, int r) { return z0[(size_t)r * J + j]; };
    auto up = [&](MargItem& it, int r) { return (int)it.values[r] < it.n_cat - 1; };
    auto lo = [&](MargItem& it, int r) { return (int)it.values[r] > 0; };
    // per-person score of rho and d loglik / d (z1, z0 | u) of each side at rho
    auto person = [&](int r, double rh, double* dla, double* dlb) -> double {
      const double s = std::sqrt(1 - rh * rh);
      if (!oa && !ob) {
        const double u = Z1(ia, r), v = Z1(ib, r), s2 = 1 - rh * rh;
        if (dla) { dla[0] = -(u - rh * v) / s2; dlb[0] = -(v - rh * u) / s2; }
        return (rh * s2 + u * v * (1 + rh * rh) - rh * (u * u + v * v)) / (s2 * s2);
      }
      if (oa != ob) {  // polyserial: c continuous, o ordinal
        const int ic = oa ? ib : ia, io = oa ? ia : ib;
        MargItem& o = oa ? a : b;
        const double u = Z1(ic, r), zz1 = Z1(io, r), zz0 = Z0(io, r);
        const double a1 = (zz1 - rh * u) / s, a0 = (zz0 - rh * u) / s;
        const double pi = std::max(Phi(a1) - Phi(a0), 1e-300);
        const double f1 = up(o, r) ? phi(a1) : 0.0, f0 = lo(o, r) ? phi(a0) : 0.0;
        if (dla) {
          double* dc = oa ? dlb : dla;
          double* dor = oa ? dla : dlb;
          dc[0] = -u - (rh / s) * (f1 - f0) / pi;  // d/du
          dor[0] = f1 / (s * pi);                   // d/dz1
          dor[1] = -f0 / (s * pi);                  // d/dz0
        }
        return (f1 * (rh * zz1 - u) - f0 * (rh * zz0 - u)) / (s * s * s * pi);
      }
      const double a1 = Z1(ia, r), a0 = Z0(ia, r), b1 = Z1(ib, r), b0 = Z0(ib, r);
      const double P = std::max(bvn_cdf(a1, b1, rh) - bvn_cdf(a0, b1, rh) - bvn_cdf(a1, b0, rh) + bvn_cdf(a0, b0, rh), 1e-300);
      const bool au = up(a, r), al = lo(a, r), bu = up(b, r), bl = lo(b, r);
      double dd = 0;
      if (au && bu) dd += bvn_pdf(a1, b1, rh);
      if (al && bu) dd -= bvn_pdf(a0, b1, rh);
      if (au && bl) dd -= bvn_pdf(a1, b0, rh);
      if (al && bl) dd += bvn_pdf(a0, b0, rh);
      if (dla) {
        auto edge = [&](double za, double zb1, double zb0) { return phi(za) * (Phi((zb1 - rh * za) / s) - Phi((zb0 - rh * za) / s)); };
        dla[0] = au ? edge(a1, b1, b0) / P : 0.0;
        dla[1] = al ? -edge(a0, b1, b0) / P : 0.0;
        dlb[0] = bu ? edge(b1, a1, a0) / P : 0.0;
        dlb[1] = bl ? -edge(b0, a1, a0) / P : 0.0;
      }
      return dd / P;
    };
    auto total = [&](double rh) {
      double part[kChunks] = {0};
      for_chunks((int)rows.size(), [&](int chunk, int lo_i, int hi_i) {
        double sum = 0;
        for (int ri = lo_i; ri < hi_i; ++ri) sum += person(rows[ri], rh, nullptr, nullptr);
        part[chunk] = sum;
      });
      double sum = 0;
      for (double v : part) sum += v;
      return sum;
    };
    double start = 0;
    if (!oa && !ob) {
      for (int r : rows) start += Z1(ia, r) * Z1(ib, r);
      start /= n;
    }
    const double rh = solve_rho(total, start);
    rho[pk] = rh;
    // scores + A21 row: sum_i score_i * d loglik2_i / d theta_1
    auto chain = [&](MargItem& it, int j, int r, const double* dl, double sc, double* row) {
      double* o = row + col_of[j];
      if (it.n_cat == 0) {  // dl[0] = d/du, u = e / sd
        const double sd = std::sqrt(it.var), u = Z1(j, r);
        o[0] += sc * (-dl[0] / sd);
        for (int k = 0; k < K; ++k) o[1 + k] += sc * X[k][r] * (-dl[0] / sd);
        o[K + 1] += sc * (-1 / (2 * it.var) + dl[0] * (-u / (2 * it.var)));
      } else {  // dl[0] = d/dz1, dl[1] = d/dz0
        const int C = it.n_cat, c = (int)it.values[r];
        if (c < C - 1) o[c] += sc * dl[0];
        if (c > 0) o[c - 1] += sc * dl[1];
        for (int k = 0; k < K; ++k) o[C - 1 + k] += sc * (-X[k][r]) * (dl[0] + dl[1]);
      }
    };
    std::vector<double> a21p((size_t)kChunks * n1, 0.0);
    for_chunks((int)rows.size(), [&](int chunk, int lo_i, int hi_i) {
      double dla[2], dlb[2];
      for (int ri = lo_i; ri < hi_i; ++ri) {
        const int r = rows[ri];
        dla[0] = dla[1] = dlb[0] = dlb[1] = 0;
        const double sc = person(r, rh, dla, dlb);
        sc0[(size_t)r * q0 + n1 + pk] = sc;
        chain(a, ia, r, dla, sc, &a21p[(size_t)chunk * n1]);
        chain(b, ib, r, dlb, sc, &a21p[(size_t)chunk * n1]);
      }
    });
    for (int chunk = 0; chunk < kChunks; ++chunk)
      for (int e = 0; e < n1; ++e) A21[pk * (size_t)n1 + e] += a21p[(size_t)chunk * n1 + e];
  }
  // B00 = sum of outer products of the score rows
  B00.assign((size_t)q0 * q0, 0.0);
  std::vector<double> bp((size_t)kChunks * q0 * q0, 0.0);
  for_chunks((int)rows.size(), [&](int chunk, int lo_i, int hi_i) {
    double* Bc = &bp[(size_t)chunk * q0 * q0];
    for (int ri = lo_i; ri < hi_i; ++ri) {
      const double* sc = &sc0[(size_t)rows[ri] * q0];
      for (int a = 0; a < q0; ++a)
        if (sc[a] != 0.0)
          for (int b = 0; b < q0; ++b) Bc[(size_t)a * q0 + b] += sc[a] * sc[b];
    }
  });
  for (int chunk = 0; chunk < kChunks; ++chunk)
    for (size_t e = 0; e < (size_t)q0 * q0; ++e) B00[e] += bp[(size_t)chunk * q0 * q0 + e];
}

}  // namespace gwsem
