// Host glue of the WLS "marginals" path: runs the SNP-independent precompute (marginals_host.cpp)
// on the people that enter the model, lays its results out in genotype-file order on the device,
// builds the statistic-vector layout shared with oracle/marginals_oracle.py, and sequences the
// per-batch kernels.
#include <algorithm>
#include <cmath>

#include "common.h"
#include "fit_program.h"
#include "kernels.h"
#include "marginals.h"
#include "model.h"

using namespace gwsem;

namespace gwsem {
void build_series_features(MargProgram& mp, int n_src, int n_pad, const std::vector<uint8_t>& inc,
                           const std::vector<double>& zz, const std::vector<uint8_t>& cat, const std::vector<double>& X,
                           const std::vector<double>& sc0, std::vector<int8_t>& packed, std::vector<double>& inv_scale,
                           std::vector<double>& tot, std::vector<double>& err);
void launch_marg_series(gwsem_engine* e, const MargProgram& mp, const int32_t* d_counts, int n_snps, double* d_summary);
size_t marg_exact_smem_bytes(const ExactProgram& ex);
void launch_marg_exact(gwsem_engine* e, const ExactProgram& ex, const uint8_t* d_rows, int64_t pitch, int n_snps, int n_src,
                       const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno, int n_pad,
                       double* d_records);
}

namespace {
// Gauss-Legendre nodes and weights on [-1, 1] (Newton on the Legendre polynomial)
void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-16) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

template <typename T>
const T* up(std::vector<DeviceBuffer<uint8_t>>& pool, const std::vector<T>& v, cudaStream_t st) {
  pool.emplace_back();
  auto& b = pool.back();
  b.ensure(std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) GW_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return reinterpret_cast<const T*>(b.p);
}
}  // namespace

void gwsem_model::prepare_marginals(const std::vector<int>& dest_of) {
  cudaStream_t st = engine->stream;
  if (marg_mode == 2) {
    // one continuous item regressed on definition variables that include the snp (and gxe products)
    const int K = (int)exo_cols.size();
    lp = LinregProgram();
    lp.K = K; lp.P = P; lp.min_variance = min_variance;
    std::vector<int> col_of_exo(K, -1);
    int ks = 0;
    for (int k = 0; k < K; ++k) {
      lp.kind[k] = exo_kind[k];
      if (exo_kind[k] != 1) { col_of_exo[k] = ks; lp.col[k] = ks++; }
    }
    lp.Ks = ks; lp.Ksp = std::max(ks, 1);
    std::vector<double> y(n_pad, 0.0), Xs((size_t)n_pad * lp.Ksp, 0.0);
    const double* yv = base[var_base[0]].data();
    for (int s = 0; s < n_src; ++s) {
      const int r = dest_of[s];
      if (r < 0 || !row_ok[r]) continue;
      y[s] = yv[r];
      for (int k = 0; k < K; ++k)
        if (exo_kind[k] != 1) Xs[(size_t)s * lp.Ksp + col_of_exo[k]] = exo_cols[k][r];
    }
    // model parameter -> position in (intercept, slopes in exo order, residual variance)
    std::vector<int32_t> par_map(P, -1);
    const int n_cells = (int)cells.size() / 5;
    for (int c = 0; c < n_cells; ++c) {
      const double* x = &cells[5 * (size_t)c];
      const int mat = (int)x[0], rr = (int)x[1], cc = (int)x[2], par = (int)x[3];
      if (par < 0) {
        if (mat == 1 && rr == 0 && cc == 0) fail("marginals: the continuous item's residual variance must be free");
        continue;
      }
      if (mat == 0) {
        int which = -1;
        for (int k = 0; k < K; ++k)
          if (exo_latent[k] == cc) which = k;
        if (rr != 0 || which < 0) fail("marginals: free path that is not a slope of the item on a definition variable");
        par_map[par] = 1 + which;
      } else if (mat == 1) {
        if (rr != 0 || cc != 0) fail("marginals: unexpected free covariance in a single-item model");
        par_map[par] = 1 + K;
      } else {
        if (cc != 0) fail("marginals: unexpected free mean in a single-item model");
        par_map[par] = 0;
      }
    }
    for (int k = 0; k < P; ++k)
      if (par_map[k] < 0) fail("marginals: parameter %d is not the mean, a slope or the variance of the item", k);
    {
      std::vector<char> seen(K + 2, 0);
      for (int k = 0; k < P; ++k) seen[par_map[k]] = 1;
      for (int k = 0; k < K + 2; ++k)
        if (!seen[k]) fail("marginals: the single-item regression must be saturated (mean, every slope and the variance free)");
    }
    lp.y = up(pool, y, st);
    lp.Xs = up(pool, Xs, st);
    lp.par_map = up(pool, par_map, st);
    lp.par_start = fp.par_start;
    return;
  }
  if (marg_mode == 1) {
    // one ordinal item regressed on snp + covariates: no-SNP probit fit as the Newton start
    std::vector<int> rows;
    for (int r = 0; r < n_rows; ++r)
      if (row_ok[r]) rows.push_back(r);
    const int C = n_categories[0];
    MargItem it;
    it.n_cat = C;
    it.values = base[var_base[0]].data();
    it.n_values = n_rows;
    for (int r : rows) {
      const double c = it.values[r];
      if (c < 0 || c >= C || c != std::floor(c)) fail("marginals: the item holds a code outside 0..%d", C - 1);
    }
    // start values: the ordered probit of the item on the plain covariates (no snp terms)
    std::vector<const double*> X;
    const int K = (int)exo_cols.size();
    std::vector<int> col_of_exo(K, -1);
    pp = ProbitProgram();
    int ks = 0;
    for (int k = 0; k < K; ++k) {
      pp.kind[k] = exo_kind[k];
      if (exo_kind[k] != 1) { col_of_exo[k] = ks; pp.col[k] = ks++; }
      if (exo_kind[k] == 0) X.push_back(exo_cols[k].data());
    }
    it.fit(rows, X);
    pp.n_cat = C; pp.K = K; pp.Kp = std::max(ks, 1); pp.P = P; pp.min_variance = min_variance;
    std::vector<uint8_t> cat(n_pad, 0);
    std::vector<double> Xd((size_t)n_pad * pp.Kp, 0.0), theta0;
    for (int s = 0; s < n_src; ++s) {
      const int r = dest_of[s];
      if (r < 0 || !row_ok[r]) continue;
      cat[s] = (uint8_t)it.values[r];
      for (int k = 0; k < K; ++k)
        if (exo_kind[k] != 1) Xd[(size_t)s * pp.Kp + col_of_exo[k]] = exo_cols[k][r];
    }
    theta0.insert(theta0.end(), it.th.begin(), it.th.end());
    for (int k = 0, plain = 0; k < K; ++k) theta0.push_back(exo_kind[k] == 0 ? it.beta[plain++] : 0.0);
    // model parameter -> position in (thresholds, slopes in the model's order); the item's variance must be 1
    std::vector<int32_t> par_map(P, -1);
    const int n_cells = (int)cells.size() / 5;
    for (int c = 0; c < n_cells; ++c) {
      const double* x = &cells[5 * (size_t)c];
      const int mat = (int)x[0], rr = (int)x[1], cc = (int)x[2], par = (int)x[3];
      if (mat == 1 && rr == 0 && cc == 0 && (par >= 0 || x[4] != 1.0)) fail("marginals: the ordinal item's variance must be fixed at 1");
      if (mat == 0 && par >= 0) {
        if (rr != 0) fail("marginals: unexpected free path in a single-item model");
        int which = -1;
        for (int k = 0; k < K; ++k)
          if (exo_latent[k] == cc) which = k;
        if (which < 0) fail("marginals: free path from a variable that is not an exogenous covariate");
        par_map[par] = C - 1 + which;
      }
      if (mat == 2 && par >= 0) fail("marginals: free means are not part of a single ordinal item model");
    }
    for (size_t t = 0; t < thresholds.size() / 4; ++t) {
      const int par = (int)thresholds[4 * t + 2];
      if (par < 0) fail("marginals: fixed thresholds are not implemented");
      par_map[par] = (int)thresholds[4 * t + 1];
    }
    for (int k = 0; k < P; ++k)
      if (par_map[k] < 0) fail("marginals: parameter %d is not a threshold or a slope of the item", k);
    pp.cat = up(pool, cat, st);
    pp.X = up(pool, Xd, st);
    pp.theta0 = up(pool, theta0, st);
    pp.par_map = up(pool, par_map, st);
    pp.par_start = fp.par_start;
    return;
  }
  // ---- the SNP-independent variables (no genotype in them) given the SNP-independent covariates: stage 1, stage 2
  // among them and their scores, once per model.  marg_mode 0 uses them as they are for every SNP without a
  // missing call; otherwise they are marg_exact_kernel's starting point.
  const int K = (int)exo_cols.size();
  std::vector<int> svars, scov, spos(p, -1), cpos(K, -1);
  for (int v = 0; v < p; ++v)
    if (kind[v] == 0) { spos[v] = (int)svars.size(); svars.push_back(v); }
  for (int k = 0; k < K; ++k)
    if (exo_kind[k] == 0) { cpos[k] = (int)scov.size(); scov.push_back(k); }
  const int J = (int)svars.size();
  std::vector<int> rows;
  for (int r = 0; r < n_rows; ++r)
    if (row_ok[r]) rows.push_back(r);
  mh = MarginalsHost();
  mh.items.resize(J);
  for (int j = 0; j < J; ++j) {
    MargItem& it = mh.items[j];
    it.n_cat = n_categories[svars[j]];
    it.values = base[var_base[svars[j]]].data();
    it.n_values = n_rows;
    if (it.n_cat > 0)
      for (int r : rows) {
        const double c = it.values[r];
        if (c < 0 || c >= it.n_cat || c != std::floor(c)) fail("marginals: manifest variable %d holds a code outside 0..%d", svars[j], it.n_cat - 1);
      }
    bool uses_any = false, uses_all = true;
    for (int k : scov) (exo_free[(size_t)svars[j] * K + k] ? uses_any : uses_all) = exo_free[(size_t)svars[j] * K + k] != 0;
    if (uses_any && !uses_all) fail("marginals: a variable regressed on some but not all exogenous covariates is not implemented");
    if (!uses_any && !scov.empty() && marg_mode == 3) fail("marginals: a data variable that is not regressed on the exogenous covariates is not implemented");
  }
  for (int k : scov) mh.X.push_back(exo_cols[k].data());
  mh.build(rows);

  // ---- theta layout shared by both kernels and the oracle: per manifest variable (continuous: intercept, slopes,
  // variance; ordinal: thresholds, slopes), then one correlation per pair (column-major lower triangle)
  ex = ExactProgram();
  ex.p = p; ex.K = K; ex.Kp = std::max(K, 1); ex.n_incl = n_incl; ex.min_variance = min_variance;
  static const bool exact_all = getenv("GWSEM_MARG_EXACT_ALL") != nullptr;   // A/B aid: every SNP through marg_exact_kernel
  ex.only_missing = marg_mode == 0 && !exact_all;
  ex.snp_exo = snp_exo >= 0;
  for (int k = 0; k < K; ++k) ex.xkind[k] = exo_kind[k];
  for (int v = 0; v < p; ++v) {
    ex.vkind[v] = kind[v] == 0 ? (n_categories[v] > 0 ? 1 : 0) : (kind[v] == 1 ? 2 : 3);
    ex.ncat[v] = n_categories[v];
    int nx = 0;
    for (int k = 0; k < K; ++k)
      if (exo_free[(size_t)v * K + k]) ex.xidx[v][nx++] = k;
    ex.nx[v] = nx;
    ex.voff[v + 1] = ex.voff[v] + (n_categories[v] > 0 ? n_categories[v] - 1 + nx : nx + 2);
  }
  ex.n1 = ex.voff[p];
  ex.npairs = p * (p - 1) / 2;
  ex.nt = ex.n1 + ex.npairs;
  auto pair_index = [&](int a, int b) {   // a > b
    int k = 0;
    for (int bb = 0; bb < b; ++bb) k += p - 1 - bb;
    return k + (a - b - 1);
  };
  {
    std::vector<double> theta0(ex.nt, 0.0);
    for (int v = 0; v < p; ++v) {
      const int o = ex.voff[v], nx = ex.nx[v];
      if (kind[v] != 0) { theta0[o + 1 + nx] = 1.0; continue; }
      const MargItem& it = mh.items[spos[v]];
      const bool uses = nx > 0;
      if (it.n_cat > 0) {
        for (int c = 0; c + 1 < it.n_cat; ++c) theta0[o + c] = it.th[c];
        for (int m = 0; m < nx; ++m) {
          const int k = ex.xidx[v][m];
          theta0[o + it.n_cat - 1 + m] = (uses && cpos[k] >= 0) ? it.beta[cpos[k]] : 0.0;
        }
      } else {
        theta0[o] = it.beta[0];
        for (int m = 0; m < nx; ++m) {
          const int k = ex.xidx[v][m];
          theta0[o + 1 + m] = (uses && cpos[k] >= 0) ? it.beta[1 + cpos[k]] : 0.0;
        }
        theta0[o + 1 + nx] = it.var;
      }
    }
    for (size_t k = 0; k < mh.pairs.size(); ++k)
      theta0[ex.n1 + pair_index(svars[mh.pairs[k].first], svars[mh.pairs[k].second])] = mh.rho[k];
    ex.theta0 = up(pool, theta0, st);
    std::vector<double> V((size_t)n_pad * p, 0.0), Xg((size_t)n_pad * ex.Kp, 0.0);
    for (int s2 = 0; s2 < n_src; ++s2) {
      const int r = dest_of[s2];
      if (r < 0 || !row_ok[r]) continue;
      for (int v = 0; v < p; ++v)
        if (var_base[v] >= 0) V[(size_t)s2 * p + v] = base[var_base[v]][r];
      for (int k = 0; k < K; ++k)
        if (exo_kind[k] != 1) Xg[(size_t)s2 * ex.Kp + k] = exo_cols[k][r];
    }
    ex.V = up(pool, V, st);
    ex.X = up(pool, Xg, st);
    std::vector<double> gl, x, w;
    const int rules[4] = {6, 12, 20, 32};
    for (int rr = 0; rr < 4; ++rr) {
      gauss_legendre(rules[rr], x, w);
      ex.gl_n[rr] = rules[rr];
      gl.insert(gl.end(), x.begin(), x.end());
      gl.insert(gl.end(), w.begin(), w.end());
    }
    ex.gl = up(pool, gl, st);
    ex.fo_theta = 3;
    ex.fo_a21 = ex.fo_theta + ex.nt;
    ex.fo_B = ex.fo_a21 + ex.npairs * ex.n1;
    ex.full_stride = ex.fo_B + ex.nt * ex.nt + 8;   // + 8 diagnostic slots (phase cycles, Newton iterations)
    ex.T = 256;
    while (ex.T > 32 && marg_exact_smem_bytes(ex) > 227 * 1024) ex.T /= 2;
    if (marg_exact_smem_bytes(ex) > 227 * 1024) fail("marginals: model too large for the per-SNP statistics kernel");
  }

  mp = MargProgram();
  mp.p = p; mp.nt = ex.nt; mp.n1t = ex.n1;
  for (int v = 0; v < p; ++v) mp.vcat[v] = n_categories[v];
  for (int v = 0; v <= p; ++v) mp.voff[v] = ex.voff[v];
  mp.full_stride = ex.full_stride; mp.fo_theta = ex.fo_theta; mp.fo_a21 = ex.fo_a21; mp.fo_B = ex.fo_B;
  mp.full_all = marg_mode == 3; mp.exact_missing = marg_mode == 0;
  if (marg_mode == 0) {
  mp.J = J; mp.K = K; mp.Kp = std::max(K, 1); mp.q0 = mh.q0; mp.q0p = mh.q0 <= 32 ? 32 : 64;
  if (mh.q0 > 64) fail("marginals: %d item-side score columns (at most 64 supported)", mh.q0);
  mp.n1 = mh.n1; mp.npii = (int)mh.pairs.size(); mp.n_incl = n_incl;
  for (int j = 0; j < J; ++j) {
    mp.n_cat[j] = mh.items[j].n_cat;
    mp.item_sd[j] = mh.items[j].n_cat == 0 ? std::sqrt(mh.items[j].var) : 1.0;
  }
  for (int j = 0; j <= J; ++j) mp.col_of[j] = mh.col_of[j];
  // person-major arrays in genotype-file order
  std::vector<double> sc0((size_t)((n_pad + 255) / 256 * 256) * mp.q0p, 0.0),  // zero rows up to whole 256-person tiles
       zz((size_t)n_pad * J * 5, 0.0), cc((size_t)n_pad * J * 3, 0.0), X((size_t)n_pad * mp.Kp, 0.0);
  std::vector<uint8_t> cat((size_t)n_pad * J, 0);
  for (int s = 0; s < n_src; ++s) {
    const int r = dest_of[s];
    if (r < 0 || !row_ok[r]) continue;
    for (int b = 0; b < mh.q0; ++b) sc0[(size_t)s * mp.q0p + b] = mh.sc0[(size_t)r * mh.q0 + b];
    for (int j = 0; j < J; ++j) {
      double* e = &zz[((size_t)s * J + j) * 5];
      e[0] = mh.z1[(size_t)r * J + j];
      e[1] = mh.z0[(size_t)r * J + j];
      if (mh.items[j].n_cat > 0) {
        e[2] = 0.3989422804014327 * std::exp(-0.5 * e[0] * e[0]);
        e[3] = 0.3989422804014327 * std::exp(-0.5 * e[1] * e[1]);
        e[4] = 0.5 * (std::erfc(-e[0] * 0.7071067811865476) - std::erfc(-e[1] * 0.7071067811865476));
        if (e[1] > 0) e[4] = 0.5 * (std::erfc(e[1] * 0.7071067811865476) - std::erfc(e[0] * 0.7071067811865476));
        if (e[0] >= 12.0) e[2] = 0.0;   // open categories: the sentinel bound carries no density
        else mp.zmax[j] = std::max(mp.zmax[j], std::fabs(e[0]));
        if (e[1] <= -12.0) e[3] = 0.0;
        else mp.zmax[j] = std::max(mp.zmax[j], std::fabs(e[1]));
        double* c = &cc[((size_t)s * J + j) * 3];
        c[0] = (e[3] - e[2]) / e[4];
        c[1] = (e[1] * e[3] - e[0] * e[2]) / e[4];
        c[2] = ((e[1] * e[1] - 1.0) * e[3] - (e[0] * e[0] - 1.0) * e[2]) / e[4];
      } else {
        cc[((size_t)s * J + j) * 3] = e[0];
      }
      if (mh.items[j].n_cat > 0) cat[(size_t)s * J + j] = (uint8_t)mh.items[j].values[r];
    }
    for (int k = 0; k < K; ++k) X[(size_t)s * mp.Kp + k] = exo_cols[k][r];
  }
  std::vector<double> theta1;
  for (int j = 0; j < J; ++j) {
    const MargItem& it = mh.items[j];
    if (it.n_cat > 0) {
      theta1.insert(theta1.end(), it.th.begin(), it.th.end());
      theta1.insert(theta1.end(), it.beta.begin(), it.beta.end());
    } else {
      theta1.insert(theta1.end(), it.beta.begin(), it.beta.end());
      theta1.push_back(it.var);
    }
  }
  // summary record layout
  const int D = 2 + J;
  int o = 3;
  mp.o_rho = o; o += J;
  mp.o_a22 = o; o += J;
  mp.o_a21snp = o; o += 2 * J;
  mp.o_a21item = o; o += mh.n1;
  mp.o_bdd = o; o += D * D;
  mp.o_bd0 = o; o += D * mh.q0;
  mp.sum_stride = o;
  mp.sc0 = up(pool, sc0, st);
  mp.zz = up(pool, zz, st);
  mp.cc = up(pool, cc, st);
  // Series pass from class sums: six products of c1..c3 per ordinal item, rounded to a 2^-s grid (|h| 2^s <= 2^30) and
  // packed as base-256 digits for the tcgen05 class-sum kernel.
  mp.cs_nf = 0;
  d_csq = nullptr;
  {
    // (a continuous item needs v and v^2 only: psi = u v, psi' = 1 - u^2 - v^2, psi'' = 6 u v; it keeps six slots)
    if (J > 0 && 6 * J <= 64) {
      const int nF = 6 * J;
      std::vector<std::vector<double>> h(nF, std::vector<double>(n_pad, 0.0));
      for (int s2 = 0; s2 < n_src; ++s2) {
        const int r = dest_of[s2];
        if (r < 0 || !row_ok[r]) continue;
        for (int j = 0; j < J; ++j) {
          const double* c = &cc[((size_t)s2 * J + j) * 3];
          double* dst[6];
          for (int k = 0; k < 6; ++k) dst[k] = &h[6 * j + k][s2];
          if (mh.items[j].n_cat > 0) {
            *dst[0] = c[0]; *dst[1] = c[1]; *dst[2] = c[0] * c[0]; *dst[3] = c[2]; *dst[4] = c[0] * c[0] * c[0]; *dst[5] = c[0] * c[1];
          } else {
            *dst[0] = c[0]; *dst[1] = c[0] * c[0];   // v, v^2
          }
        }
      }
      std::vector<std::vector<int64_t>> q(nF, std::vector<int64_t>(n_pad, 0));
      std::vector<double> inv(nF), tot(nF, 0.0);
      for (int f = 0; f < nF; ++f) {
        double mx = 0.0;
        for (double v : h[f]) mx = std::max(mx, std::fabs(v));
        int ex = 0;
        if (mx > 0) std::frexp(mx, &ex);
        const double sc = std::ldexp(1.0, 30 - ex);
        inv[f] = std::ldexp(1.0, ex - 30);
        long double acc = 0;
        for (int s2 = 0; s2 < n_pad; ++s2) {
          q[f][s2] = std::llrint(h[f][s2] * sc);
          acc += (long double)q[f][s2];
        }
        tot[f] = (double)acc * inv[f];
      }
      std::vector<uint8_t> inc(n_pad, 0), kept(n_pad, 0);
      for (int s2 = 0; s2 < n_src; ++s2) {
        const int r = dest_of[s2];
        inc[s2] = r >= 0 && row_ok[r];
        kept[s2] = r >= 0;
      }
      std::vector<int8_t> packed;
      tc5_pack_features(q, n_pad, inc, kept, packed);
      d_csq = up(pool, packed, st);
      d_cs_inv_scale = up(pool, inv, st);
      mp.cs_tot = up(pool, tot, st);
      mp.cs_nf = nF;
    }
  }
  // Series path (marg_series.cu): every SNP-dependent statistic of a call-complete hardcall SNP from genotype-class sums
  // of per-person series coefficients (all items ordinal; GWSEM_NO_SERIES=1 switches it off, an A/B aid)
  mp.se_nf = 0;
  d_seq = nullptr;
  {
    bool all_ordinal = J > 0;
    for (int j = 0; j < J; ++j) all_ordinal = all_ordinal && mh.items[j].n_cat > 0;
    const bool off = getenv("GWSEM_NO_SERIES") != nullptr;   // read per model: the tests build both variants in one process
    if (all_ordinal && !off && mh.q0 <= 64) {
      std::vector<uint8_t> inc(n_pad, 0);
      for (int s2 = 0; s2 < n_src; ++s2) {
        const int r = dest_of[s2];
        inc[s2] = r >= 0 && row_ok[r];
      }
      std::vector<int8_t> packed;
      std::vector<double> inv, tot, err;
      build_series_features(mp, n_src, n_pad, inc, zz, cat, X, sc0, packed, inv, tot, err);
      if (mp.se_nf > 0) {   // (0: a feature too heavy-tailed for three digits; the per-person kernel serves the model)
        d_seq = up(pool, packed, st);
        d_se_inv_scale = up(pool, inv, st);
        mp.se_tot = up(pool, tot, st);
        mp.se_err = up(pool, err, st);
      }
    }
  }
  mp.cat = up(pool, cat, st);
  mp.X = up(pool, X, st);
  mp.B00 = up(pool, mh.B00, st);
  mp.A21ii = up(pool, mh.A21, st);
  mp.theta1 = up(pool, theta1, st);
  mp.rho_ii = up(pool, mh.rho, st);
  {
    // item-side block of Muthen's A (block-diagonal stage-1 information, polychoric rows A21 | diagonal) -> W = A_cc^-1
    // and acov_cc = W B00 W', the SNP-independent part of acov(theta) (marg_fit_fast_kernel)
    const int q = mh.q0, n1 = mh.n1;
    std::vector<double> Acc((size_t)q * q, 0.0), W((size_t)q * q, 0.0), T((size_t)q * q, 0.0), AC((size_t)q * q, 0.0);
    auto item_of = [&](int c) { int j = 0; while (j + 1 < J && c >= mh.col_of[j + 1]) ++j; return j; };
    for (int x = 0; x < q; ++x)
      for (int y = 0; y < q; ++y) {
        double a = 0.0;
        if (x < n1 && y < n1) { if (item_of(x) == item_of(y)) a = mh.B00[(size_t)x * q + y]; }
        else if (x >= n1) {
          if (y < n1) a = mh.A21[(size_t)(x - n1) * n1 + y];
          else if (y == x) a = mh.B00[(size_t)x * q + x];
        }
        Acc[(size_t)x * q + y] = a;
        W[(size_t)x * q + y] = x == y ? 1.0 : 0.0;
      }
    bool ok = true;
    for (int c = 0; c < q && ok; ++c) {   // Gauss-Jordan with partial pivoting
      int piv = c;
      for (int r = c + 1; r < q; ++r)
        if (std::fabs(Acc[(size_t)r * q + c]) > std::fabs(Acc[(size_t)piv * q + c])) piv = r;
      if (!(std::fabs(Acc[(size_t)piv * q + c]) > 0.0)) { ok = false; break; }
      if (piv != c)
        for (int k = 0; k < q; ++k) { std::swap(Acc[(size_t)c * q + k], Acc[(size_t)piv * q + k]); std::swap(W[(size_t)c * q + k], W[(size_t)piv * q + k]); }
      const double inv = 1.0 / Acc[(size_t)c * q + c];
      for (int k = 0; k < q; ++k) { Acc[(size_t)c * q + k] *= inv; W[(size_t)c * q + k] *= inv; }
      for (int r = 0; r < q; ++r) {
        if (r == c) continue;
        const double f = Acc[(size_t)r * q + c];
        if (f == 0.0) continue;
        for (int k = 0; k < q; ++k) { Acc[(size_t)r * q + k] -= f * Acc[(size_t)c * q + k]; W[(size_t)r * q + k] -= f * W[(size_t)c * q + k]; }
      }
    }
    mp.Wcc = nullptr;
    mp.acov_cc = nullptr;
    if (ok) {
      for (int x = 0; x < q; ++x)
        for (int y = 0; y < q; ++y) {
          double v = 0.0;
          for (int k = 0; k < q; ++k) v += W[(size_t)x * q + k] * mh.B00[(size_t)k * q + y];
          T[(size_t)x * q + y] = v;
        }
      for (int x = 0; x < q; ++x)
        for (int y = 0; y <= x; ++y) {
          double v = 0.0;
          for (int k = 0; k < q; ++k) v += T[(size_t)x * q + k] * W[(size_t)y * q + k];
          AC[(size_t)x * q + y] = AC[(size_t)y * q + x] = v;
        }
      mp.Wcc = up(pool, W, st);
      mp.acov_cc = up(pool, AC, st);
    }
  }
  {
    std::vector<double> h11;
    for (int j = 0; j < J; ++j) {
      mp.h11_of[j] = (int)h11.size();
      h11.insert(h11.end(), mh.items[j].hess.begin(), mh.items[j].hess.end());
    }
    mp.h11_of[J] = (int)h11.size();
    mp.H11 = up(pool, h11, st);
  }
  }  // marg_mode == 0
  // statistic vector: (mean | thresholds per variable), slopes, variances, covariances
  std::vector<int32_t> kind_, a_, b_, th_, tpar;
  std::vector<double> tval;
  auto push = [&](int kd, int a, int b, int th) { kind_.push_back(kd); a_.push_back(a); b_.push_back(b); th_.push_back(th); tpar.push_back(-1); tval.push_back(0.0); };
  for (int v = 0; v < p; ++v) {
    if (n_categories[v] == 0) push(0, v, 0, ex.voff[v]);
    else
      for (int c = 0; c + 1 < n_categories[v]; ++c) {
        push(1, v, c, ex.voff[v] + c);
        bool found = false;
        for (size_t t = 0; t < thresholds.size() / 4; ++t)
          if ((int)thresholds[4 * t] == v && (int)thresholds[4 * t + 1] == c) {
            tpar.back() = (int)thresholds[4 * t + 2];
            tval.back() = thresholds[4 * t + 3];
            found = true;
          }
        if (!found) fail("marginals: no threshold %d given for ordinal variable %d", c + 1, v);
      }
  }
  for (int v = 0; v < p; ++v)
    for (int m = 0; m < ex.nx[v]; ++m)
      push(2, v, ex.xidx[v][m], ex.voff[v] + (n_categories[v] > 0 ? n_categories[v] - 1 : 1) + m);
  for (int v = 0; v < p; ++v)
    if (n_categories[v] == 0) push(3, v, v, ex.voff[v + 1] - 1);
  for (int b = 0; b < p; ++b)
    for (int a = b + 1; a < p; ++a) push(4, a, b, ex.n1 + pair_index(a, b));
  mp.n_stat = (int)kind_.size();
  if (mp.n_stat > 128 || ex.nt > 128) fail("marginals: model too large");
  mp.stat_kind = up(pool, kind_, st);
  mp.stat_a = up(pool, a_, st);
  mp.stat_b = up(pool, b_, st);
  mp.stat_theta = up(pool, th_, st);
  fp.thr_par = up(pool, tpar, st);
  fp.thr_val = up(pool, tval, st);
  fp.exo_latent = up(pool, exo_latent, st);

  // Parameters that enter exactly one statistic (and are not bounded) leave the iteration: read
  // the structural zero pattern of the Jacobian from two generic points.
  {
    const int ns = mp.n_stat;
    DeviceBuffer<double> dj;
    dj.ensure((size_t)2 * ns * P);
    mp.n_a = 0;
    launch_marg_pattern(engine, fp, mp, dj.p);
    std::vector<double> jj((size_t)2 * ns * P);
    GW_CUDA(cudaMemcpyAsync(jj.data(), dj.p, jj.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> a_par, a_row, b_par, b_row;
    std::vector<char> row_taken(ns, 0), par_gone(P, 0);
    for (int k = 0; k < P; ++k) {
      if (par_lbound[k] == par_lbound[k]) continue;
      int only = -1, cnt = 0;
      for (int e = 0; e < ns; ++e)
        if (jj[(size_t)e * P + k] != 0.0 || jj[(size_t)(ns + e) * P + k] != 0.0) { only = e; ++cnt; }
      if (cnt == 1 && !row_taken[only]) { row_taken[only] = 1; par_gone[k] = 1; a_par.push_back(k); a_row.push_back(only); }
    }
    for (int e = 0; e < ns; ++e) if (!row_taken[e]) b_row.push_back(e);
    for (int k = 0; k < P; ++k) if (!par_gone[k]) b_par.push_back(k);
    mp.n_init = 0;
    for (int e : b_row) {
      if (kind_[e] != 0 && kind_[e] != 3) continue;
      for (int k : b_par) {
        if (par_lbound[k] == par_lbound[k] && kind_[e] == 0) continue;
        if (jj[(size_t)e * P + k] == 1.0 && jj[(size_t)(ns + e) * P + k] == 1.0 && mp.n_init < 4) {
          mp.init_par[mp.n_init] = k;
          mp.init_row[mp.n_init] = e;
          ++mp.n_init;
          break;
        }
      }
    }
    if (!a_par.empty()) {
      mp.n_a = (int)a_par.size(); mp.n_b = (int)b_row.size(); mp.n_beta = (int)b_par.size();
      mp.elim_a_par = up(pool, a_par, st); mp.elim_a_row = up(pool, a_row, st);
      mp.elim_b_par = up(pool, b_par, st); mp.elim_b_row = up(pool, b_row, st);
    }
  }

  // Warm start.  The reference restarts every SNP from the model's start values so that a row does
  // not depend on its neighbours (R/model.R:188); the same holds here with a start that is fixed
  // per model: the estimates for a reference genotype column (codes 0,1,2,0,1,2,... down the file,
  // unrelated to everybody's phenotype), fitted once from the model's own start values.  The item
  // side of the solution barely moves between SNPs, so the per-SNP iteration starts next to it.
  {
    const int64_t rb = (n_pad / 4 + 15) / 16 * 16;
    std::vector<uint8_t> ref(rb, 0);
    for (int i = 0; i < n_src; ++i) ref[i >> 2] |= (uint8_t)((i % 3) << (2 * (i & 3)));
    const uint8_t* d_ref = up(pool, ref, st);
    DeviceBuffer<double> dd;
    DeviceBuffer<int32_t> di;
    dd.ensure((size_t)P + (size_t)P * P + 2);
    di.ensure(5);
    gwsem_result r;
    r.est = dd.p; r.vcov = dd.p + P; r.fit = dd.p + P + (size_t)P * P; r.maf = r.fit + 1;
    r.n_used = di.p; r.status = di.p + 1; r.catch_code = di.p + 2; r.catch_var = di.p + 3; r.iterations = di.p + 4;
    fp.par_warm = nullptr;
    run_marginals(d_ref, rb, nullptr, nullptr, 1, 0, r);
    int32_t flags[5];
    std::vector<double> warm(P);
    GW_CUDA(cudaMemcpyAsync(flags, di.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaMemcpyAsync(warm.data(), dd.p, P * sizeof(double), cudaMemcpyDeviceToHost, st));
    GW_CUDA(cudaStreamSynchronize(st));
    if (flags[1] == GWSEM_STATUS_OK && flags[2] == GWSEM_CATCH_NONE) fp.par_warm = up(pool, warm, st);
  }
}

// GWSEM_EXACT_DIAG=1 (development aid): mean cycles per phase and Newton iterations of marg_exact_kernel, on stderr
void gwsem_model::exact_diag(int n) {
  static const bool on = getenv("GWSEM_EXACT_DIAG") != nullptr;
  if (!on) return;
  std::vector<double> h((size_t)n * ex.full_stride);
  GW_CUDA(cudaMemcpyAsync(h.data(), d_full.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, engine->stream));
  GW_CUDA(cudaStreamSynchronize(engine->stream));
  double d[8] = {0};
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    const double* rec = &h[(size_t)i * ex.full_stride];
    if (rec[0] < 0 || rec[1] != 0.0) continue;
    ++cnt;
    for (int e = 0; e < 8; ++e) d[e] += rec[ex.fo_B + ex.nt * ex.nt + e];
  }
  if (cnt)
    fprintf(stderr, "marg_exact: %d SNPs, mean Mcycles A %.2f B %.2f C %.2f D %.2f, iterations B %.1f C %.1f (T = %d)\n", cnt,
            d[0] / cnt / 1e6, d[1] / cnt / 1e6, d[2] / cnt / 1e6, d[3] / cnt / 1e6, d[4] / cnt, d[5] / cnt, ex.T);
}

void gwsem_model::run_marginals(const uint8_t* d_packed, int64_t pitch, const double* d_geno, const double* d_maf_in,
                                int n, int plink1, const gwsem_result& r) {
  if (marg_mode == 1) {
    const double* maf1 = d_maf_in;
    if (!d_geno) {
      d_counts.ensure(8 * (size_t)n);
      d_miss.ensure(16);
      d_mafs.ensure(n);
      launch_count_missing(engine, d_packed, pitch, n, plink1, layout, d_featT, 0, 1, d_counts.p, d_miss.p);
      launch_maf_from_counts(engine, d_counts.p, layout.n_kept, n, d_mafs.p);
      maf1 = d_mafs.p;
    }
    launch_probit_snp(engine, pp, d_packed, pitch, n, n_src, layout.d_inc8, plink1, d_geno, n_pad, maf1, r);
    return;
  }
  if (marg_mode == 2) {
    const double* maf1 = d_maf_in;
    if (!d_geno) {
      d_counts.ensure(8 * (size_t)n);
      d_miss.ensure(16);
      d_mafs.ensure(n);
      launch_count_missing(engine, d_packed, pitch, n, plink1, layout, d_featT, 0, 1, d_counts.p, d_miss.p);
      launch_maf_from_counts(engine, d_counts.p, layout.n_kept, n, d_mafs.p);
      maf1 = d_mafs.p;
    }
    launch_linreg_snp(engine, lp, d_packed, pitch, n, n_src, layout.d_inc8, plink1, d_geno, n_pad, maf1, r);
    return;
  }
  const double* maf = d_maf_in;
  if (!d_geno) {
    d_counts.ensure(8 * (size_t)n);
    d_miss.ensure(16);
    d_mafs.ensure(n);
    launch_count_missing(engine, d_packed, pitch, n, plink1, layout, d_featT, 0, 1, d_counts.p, d_miss.p);
    launch_maf_from_counts(engine, d_counts.p, layout.n_kept, n, d_mafs.p);
    maf = d_mafs.p;
  }
  MargProgram mpl = mp;
  d_full.ensure((size_t)n * ex.full_stride);
  mpl.full = d_full.p;
  if (marg_mode == 3) {
    // nothing is SNP-independent: every statistic of every SNP from marg_exact_kernel
    launch_marg_exact(engine, ex, d_packed, pitch, n, n_src, layout.d_inc8, d_geno ? nullptr : d_counts.p, plink1, d_geno,
                      n_pad, d_full.p);
    exact_diag(n);
    launch_marg_fit(engine, fp, mpl, n, nullptr, maf, r);
    return;
  }
  d_summary.ensure((size_t)n * mp.sum_stride);
  mpl.cs_sums = nullptr;
  // (with the series path on, the per-person kernel only sees the SNPs that path turned down, nearly all of them with a
  // series root to start from: the six-product class sums for its iteration 0 are not worth a launch over the batch)
  if (!d_geno && mp.cs_nf > 0 && mp.se_nf == 0) {
    d_cs_sums.ensure((size_t)n * 2 * mp.cs_nf);
    launch_class_sums_tc5(engine, d_packed, pitch, n, plink1, layout, d_csq, d_cs_inv_scale, mp.cs_nf, d_cs_sums.p, nullptr);
    mpl.cs_sums = d_cs_sums.p;
  }
  // SNPs with a missing call: the people retained differ from the model's, so every statistic is recomputed over them
  // (marg_exact_kernel; the others leave at once).  It goes first, on a side stream: its few long-running CTAs keep
  // their SMs while marg_stats_kernel streams the whole batch through the rest.
  if (!side_stream) {
    GW_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    GW_CUDA(cudaEventCreateWithFlags(&ev_side_go, cudaEventDisableTiming));
    GW_CUDA(cudaEventCreateWithFlags(&ev_side_done, cudaEventDisableTiming));
  }
  {
    cudaStream_t main_stream = engine->stream;
    GW_CUDA(cudaEventRecord(ev_side_go, main_stream));
    GW_CUDA(cudaStreamWaitEvent(side_stream, ev_side_go, 0));
    engine->stream = side_stream;
    try {
      launch_marg_exact(engine, ex, d_packed, pitch, n, n_src, layout.d_inc8, d_geno ? nullptr : d_counts.p, plink1, d_geno,
                        n_pad, d_full.p);
    } catch (...) {
      engine->stream = main_stream;
      throw;
    }
    engine->stream = main_stream;
    GW_CUDA(cudaEventRecord(ev_side_done, side_stream));
  }
  mpl.se_done = nullptr;
  if (!d_geno && mp.se_nf > 0) {
    // call-complete SNPs with a small correlation: the whole record from class sums, no pass over people
    d_se_sums.ensure((size_t)n * 2 * mp.se_nf);
    d_se_done.ensure(n);
    launch_class_sums_tc5_wide(engine, d_packed, pitch, n, plink1, layout, d_seq, d_se_inv_scale, mp.se_n_wide, mp.se_n_narrow,
                               d_se_sums.p);
    mpl.se_sums = d_se_sums.p;
    mpl.se_done = d_se_done.p;
    launch_marg_series(engine, mpl, d_counts.p, n, d_summary.p);
  }
  launch_marg_stats(engine, mpl, d_packed, pitch, n, n_src, layout.d_inc8, d_geno ? nullptr : d_counts.p, plink1, d_geno,
                    n_pad, d_summary.p);
  GW_CUDA(cudaStreamWaitEvent(engine->stream, ev_side_done, 0));
  exact_diag(n);
  static const bool dbg = getenv("GWSEM_DEBUG_SYNC") != nullptr;   // development aid: attribute a device fault to its kernel
  if (dbg) {
    GW_CUDA(cudaStreamSynchronize(engine->stream));
    fprintf(stderr, "gwsem debug: statistics kernels of %d SNPs done\n", n);
  }
  launch_marg_fit(engine, fp, mpl, n, d_summary.p, maf, r);
  if (dbg) {
    GW_CUDA(cudaStreamSynchronize(engine->stream));
    fprintf(stderr, "gwsem debug: marg_fit of %d SNPs done\n", n);
  }
}
