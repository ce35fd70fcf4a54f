// ML / FIML for continuous manifest variables (reference call sites
// R/model.R:6 mxFitFunctionML, 166-173 GradientDescent + NumericDeriv + StandardError, 403-418
// exogenous covariates as definition variables).  One warp per SNP.
//
//   -2 lnL = sum_g [ n_g (p_g log 2 pi + log |Sigma_g|) + tr(Sigma_g^-1 T_g) ]
//   T_g = Myy - Mwy' C' - C R ,  R = Mwy - Mww C' ,  C = [mu, slopes] restricted to group g
// over two groups of people: genotype observed (all manifest variables) and genotype missing (the
// SNP-independent manifest variables only).  The augmented moments M of [1, covariates, manifest]
// come from the raw moment table of the statistics kernels (stats.cu), so nothing here touches a
// person.  Gradient analytically (RAM derivatives contracted with Q = n Sigma^-1 - Sigma^-1 T
// Sigma^-1 and U = Sigma^-1 R'), Newton steps on the finite-difference Hessian of that gradient
// (damped when it is not positive definite), parameter covariance 2 H^-1 with H by central
// differences at the estimate.  Mirrors oracle/ml_oracle.py.
#include "common.h"
#include "fit_device.cuh"
#include "kernels.h"

namespace gwsem {

namespace {

constexpr int LANES = 32;
constexpr double kLog2Pi = 1.8378770664093453;

struct MlScratch {
  int A, S, Z, C, T, sig, Mv, mu, Eq, Eu, V1, VU, WA, WS, WU, Sg, G, Tm, GT, Cf, R, U, mom, theta, cand, g, g1, g2, delta, H, Hc, gi, total;
};

__host__ __device__ inline MlScratch ml_layout(int p, int nv, int P, int dw) {
  MlScratch m;
  int o = 0;
  auto take = [&](int n) { int at = o; o += n; return at; };
  m.A = take(nv * nv); m.S = take(nv * nv); m.Z = take(nv * nv); m.C = take(nv * nv); m.T = take(nv * nv);
  m.sig = take(p * (p + 1) / 2);
  m.Mv = take(nv); m.mu = take(nv);
  m.Eq = take(p * p); m.Eu = take(p * dw);
  m.V1 = take(nv * p); m.VU = take(nv * dw);
  m.WA = take(nv * nv); m.WS = take(nv * nv); m.WU = take(nv * nv);
  m.Sg = take(p * p); m.G = take(p * p); m.Tm = take(p * p); m.GT = take(p * p);
  m.Cf = take(p * dw); m.R = take(dw * p); m.U = take(p * dw);
  m.mom = take(2 * (dw * dw + dw * p + p * p));
  m.theta = take(P); m.cand = take(P); m.g = take(P); m.g1 = take(P); m.g2 = take(P); m.delta = take(P);
  m.H = take(P * P); m.Hc = take(P * P);
  m.gi = take(p);  // group's manifest indices, as doubles
  m.total = o;
  return m;
}

struct MlCtx {
  const FitProgram& fp;
  const MlScratch& ms;
  double* sm;
  int lane, p, nv, P, dw;
  double n_g[2];
  int pg[2];
};

// -2 lnL at thv; gradient into gout (if not null).  Returns +inf when a group's covariance is not
// positive definite.
__device__ double ml_eval(MlCtx& cx, const double* thv, double* gout) {
  const FitProgram& fp = cx.fp;
  const MlScratch& ms = cx.ms;
  double* sm = cx.sm;
  const int lane = cx.lane, p = cx.p, nv = cx.nv, dw = cx.dw, K = dw - 1;
  WarpScratch ws{};
  ws.A = ms.A; ws.S = ms.S; ws.Z = ms.Z; ws.C = ms.C; ws.T = ms.T; ws.sig = ms.sig;
  ModelEval<LANES> ev{fp, ws, sm, lane};
  ev.implied(thv, sm + ms.sig);
  double *Z = sm + ms.Z, *C = sm + ms.C, *Mv = sm + ms.Mv, *mu = sm + ms.mu;
  for (int e = lane; e < nv; e += LANES) Mv[e] = 0.0;
  gsync<LANES>();
  for (int e = lane; e < fp.n_cells; e += LANES)
    if (fp.cell_mat[e] == 2) Mv[fp.cell_col[e]] = fp.cell_par[e] >= 0 ? thv[fp.cell_par[e]] : fp.cell_val[e];
  gsync<LANES>();
  for (int i = lane; i < nv; i += LANES) {
    double v = 0.0;
    for (int k = 0; k < nv; ++k) v += Z[i * nv + k] * Mv[k];
    mu[i] = v;
  }
  double *Eq = sm + ms.Eq, *Eu = sm + ms.Eu;
  for (int e = lane; e < p * p; e += LANES) Eq[e] = 0.0;
  for (int e = lane; e < p * dw; e += LANES) Eu[e] = 0.0;
  gsync<LANES>();
  double f = 0.0;
  double *Sg = sm + ms.Sg, *G = sm + ms.G, *Tm = sm + ms.Tm, *GT = sm + ms.GT, *Cf = sm + ms.Cf, *R = sm + ms.R, *U = sm + ms.U;
  double* gi = sm + ms.gi;
  for (int g = 0; g < 2; ++g) {
    if (!(cx.n_g[g] > 0.0)) continue;
    const int pg = cx.pg[g];
    if (pg == 0) continue;
    // manifest indices of the group
    if (lane == 0) {
      int c = 0;
      for (int i = 0; i < p; ++i)
        if (g == 0 || fp.man_static[i]) gi[c++] = i;
    }
    gsync<LANES>();
    const double* mom = sm + ms.mom + g * (dw * dw + dw * p + p * p);
    const double *Mww = mom, *Mwy = mom + dw * dw, *Myy = mom + dw * dw + dw * p;
    for (int e = lane; e < pg * pg; e += LANES) Sg[e] = C[(int)gi[e / pg] * nv + (int)gi[e % pg]];
    for (int e = lane; e < pg * dw; e += LANES) {
      const int a = e / dw, m = e % dw, i = (int)gi[a];
      Cf[e] = m == 0 ? mu[i] : Z[i * nv + fp.exo_latent[m - 1]];
    }
    gsync<LANES>();
    if (!warp_cholesky<LANES>(Sg, pg, lane)) return INFINITY;
    double logdet = 0.0;
    for (int a = 0; a < pg; ++a) logdet += 2.0 * log(Sg[a * pg + a]);
    // G = Sigma^-1, column per lane
    for (int c = lane; c < pg; c += LANES) {
      double* x = G + c;
      for (int i = 0; i < pg; ++i) {
        double v = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < i; ++k) v -= Sg[i * pg + k] * x[k * pg];
        x[i * pg] = v / Sg[i * pg + i];
      }
      for (int i = pg - 1; i >= 0; --i) {
        double v = x[i * pg];
        for (int k = i + 1; k < pg; ++k) v -= Sg[k * pg + i] * x[k * pg];
        x[i * pg] = v / Sg[i * pg + i];
      }
    }
    // R[m][a] = Mwy[m][i_a] - sum_l Mww[m][l] Cf[a][l]
    for (int e = lane; e < dw * pg; e += LANES) {
      const int m = e / pg, a = e % pg;
      double v = Mwy[m * p + (int)gi[a]];
      for (int l = 0; l < dw; ++l) v -= Mww[m * dw + l] * Cf[a * dw + l];
      R[m * pg + a] = v;
    }
    gsync<LANES>();
    // T[a][b] = Myy[a][b] - sum_m Mwy[m][a] Cf[b][m] - sum_m Cf[a][m] R[m][b]
    for (int e = lane; e < pg * pg; e += LANES) {
      const int a = e / pg, b = e % pg;
      double v = Myy[(int)gi[a] * p + (int)gi[b]];
      for (int m = 0; m < dw; ++m) v -= Mwy[m * p + (int)gi[a]] * Cf[b * dw + m] + Cf[a * dw + m] * R[m * pg + b];
      Tm[e] = v;
    }
    gsync<LANES>();
    double tr = 0.0;
    for (int e = lane; e < pg * pg; e += LANES) tr += G[e] * Tm[e];
#pragma unroll
    for (int o = 16; o; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
    f += cx.n_g[g] * (pg * kLog2Pi + logdet) + tr;
    if (gout) {
      for (int e = lane; e < pg * pg; e += LANES) {
        const int a = e / pg, b = e % pg;
        double v = 0.0;
        for (int k = 0; k < pg; ++k) v += G[a * pg + k] * Tm[k * pg + b];
        GT[e] = v;
      }
      for (int e = lane; e < pg * dw; e += LANES) {
        const int a = e / dw, m = e % dw;
        double v = 0.0;
        for (int b = 0; b < pg; ++b) v += G[a * pg + b] * R[m * pg + b];
        U[e] = v;
      }
      gsync<LANES>();
      for (int e = lane; e < pg * pg; e += LANES) {
        const int a = e / pg, b = e % pg;
        double v = cx.n_g[g] * G[e];
        for (int k = 0; k < pg; ++k) v -= GT[a * pg + k] * G[k * pg + b];
        Eq[(int)gi[a] * p + (int)gi[b]] += v;
      }
      for (int e = lane; e < pg * dw; e += LANES) Eu[(int)gi[e / dw] * dw + e % dw] += U[e];
      gsync<LANES>();
    }
  }
  if (!gout) return f;
  // tables: V1 = Z[:p,:]' Eq, VU = Z[:p,:]' Eu ; WA = V1 C[:, :p]', WS = V1 Z[:p, :], WU = VU [mu; Z[:, exo]]'
  double *V1 = sm + ms.V1, *VU = sm + ms.VU, *WA = sm + ms.WA, *WS = sm + ms.WS, *WU = sm + ms.WU;
  for (int e = lane; e < nv * p; e += LANES) {
    const int r = e / p, j = e % p;
    double v = 0.0;
    for (int i = 0; i < p; ++i) v += Z[i * nv + r] * Eq[i * p + j];
    V1[e] = v;
  }
  for (int e = lane; e < nv * dw; e += LANES) {
    const int r = e / dw, m = e % dw;
    double v = 0.0;
    for (int i = 0; i < p; ++i) v += Z[i * nv + r] * Eu[i * dw + m];
    VU[e] = v;
  }
  gsync<LANES>();
  for (int e = lane; e < nv * nv; e += LANES) {
    const int r = e / nv, c = e % nv;
    double va = 0.0, vs = 0.0;
    for (int j = 0; j < p; ++j) { va += V1[r * p + j] * C[c * nv + j]; vs += V1[r * p + j] * Z[j * nv + c]; }
    double vu = VU[r * dw] * mu[c];
    for (int k = 0; k < K; ++k) vu += VU[r * dw + 1 + k] * Z[c * nv + fp.exo_latent[k]];
    WA[e] = va; WS[e] = vs; WU[e] = vu;
  }
  gsync<LANES>();
  for (int k = lane; k < cx.P; k += LANES) {
    double v = 0.0;
    for (int ce = fp.par_cell_ptr[k]; ce < fp.par_cell_ptr[k + 1]; ++ce) {
      const int id = fp.par_cells[ce], m = fp.cell_mat[id], r = fp.cell_row[id], c = fp.cell_col[id];
      if (m == 0) v += 2.0 * (WA[r * nv + c] - WU[r * nv + c]);
      else if (m == 1) v += WS[r * nv + c];
      else v -= 2.0 * VU[c * dw];
    }
    gout[k] = v;
  }
  gsync<LANES>();
  return f;
}

__global__ void __launch_bounds__(32)
    ml_fit_kernel(FitProgram fp, int n_snps, const double* __restrict__ Rtab, const int32_t* __restrict__ n_rows,
                  const double* __restrict__ maf_in, gwsem_result res) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x, snp = blockIdx.x;
  if (snp >= n_snps) return;
  const int p = fp.p, nv = fp.nv, P = fp.P, K = fp.n_exo, dw = 1 + K, pa = p + K, p1 = pa + 1;
  const MlScratch ms = ml_layout(p, nv, P, dw);
  double* sm = smem_d;
  const double* Rs = Rtab + (int64_t)snp * 5 * fp.nf;
  const int n = n_rows[snp];
  double* est = res.est + (int64_t)snp * P;
  double* vcov = res.vcov + (int64_t)snp * P * P;
  const double na = __longlong_as_double((long long)GWSEM_NA_BITS);
  for (int k = lane; k < P; k += LANES) est[k] = fp.par_start[k];
  for (int k = lane; k < P * P; k += LANES) vcov[k] = na;
  if (lane == 0) {
    res.maf[snp] = maf_in[snp];
    res.n_used[snp] = n;
    res.fit[snp] = na;
    res.status[snp] = GWSEM_STATUS_NA;
    res.catch_code[snp] = GWSEM_CATCH_NONE;
    res.catch_var[snp] = -1;
    res.iterations[snp] = 0;
  }
  // ---- augmented moments.  Moment variable u: manifest i -> i, covariate k -> p + k, constant -> pa.
  auto moment = [&](int u, int v, int g) -> double {
    if (u > v) { const int t = u; u = v; v = t; }
    if (u == pa) return g == 0 ? (double)n : (double)(fp.n_incl - n);
    const int idx = fp.idx4[u * p1 + v];
    if (g == 0) return Rs[idx];
    return idx < fp.nf ? fp.tot[idx] - Rs[idx] : 0.0;   // missing-call people: SNP-independent moments only
  };
  auto wvar = [&](int m) { return m == 0 ? pa : p + (m - 1); };
  double* mom = sm + ms.mom;
  for (int g = 0; g < 2; ++g) {
    double* base = mom + g * (dw * dw + dw * p + p * p);
    for (int e = lane; e < dw * dw; e += LANES) base[e] = moment(wvar(e / dw), wvar(e % dw), g);
    for (int e = lane; e < dw * p; e += LANES) base[dw * dw + e] = moment(wvar(e / p), e % p, g);
    for (int e = lane; e < p * p; e += LANES) base[dw * dw + dw * p + e] = moment(e / p, e % p, g);
  }
  gsync<LANES>();
  MlCtx cx{fp, ms, sm, lane, p, nv, P, dw, {0.0, 0.0}, {0, 0}};
  cx.n_g[0] = (double)n;
  cx.pg[0] = p;
  int ps = 0;
  for (int i = 0; i < p; ++i) ps += fp.man_static[i] ? 1 : 0;
  cx.pg[1] = ps;
  cx.n_g[1] = (fp.snp_exo || ps == 0) ? 0.0 : (double)(fp.n_incl - n);
  // ---- observed variances (mxData minVariance); an exogenous snp is checked as covariate
  if (n < 2) {
    if (lane == 0) { res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE; res.catch_var[snp] = fp.snp_exo ? -2 : 0; }
    return;
  }
  {
    const double* b0 = mom;
    int bad = -1;
    for (int i = 0; i < p && bad < 0; ++i) {
      const double m1 = b0[dw * dw + i] / n, m2 = b0[dw * dw + dw * p + i * p + i] / n;
      if (!((m2 - m1 * m1) * (double)n / (double)(n - 1) >= fp.min_variance)) bad = i;
    }
    if (bad < 0 && fp.snp_exo) {
      const int m = 1 + fp.snp_exo_index;
      const double m1 = b0[m] / n, m2 = b0[m * dw + m] / n;
      if (!((m2 - m1 * m1) * (double)n / (double)(n - 1) >= fp.min_variance)) bad = -2;
    }
    if (bad != -1) {
      if (lane == 0) { res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE; res.catch_var[snp] = bad; }
      return;
    }
  }
  // ---- Newton iterations
  double *theta = sm + ms.theta, *cand = sm + ms.cand, *g = sm + ms.g, *g1 = sm + ms.g1, *g2 = sm + ms.g2;
  double *delta = sm + ms.delta, *H = sm + ms.H, *Hc = sm + ms.Hc;
  // Start: the per-model warm start (or the model's start values), with the free mean and variance
  // of every SNP-dependent manifest variable that nothing points at moved to this SNP's sample
  // moments: they are what changes most from SNP to SNP (allele frequency).
  auto load_start = [&](const double* from) {
    for (int k = lane; k < P; k += LANES) theta[k] = from[k];
    gsync<LANES>();
    if (lane == 0) {
      const double* b0 = mom;
      for (int v = 0; v < p; ++v) {
        if (fp.man_static[v]) continue;
        bool target = false;
        for (int e = 0; e < fp.n_cells; ++e)
          if (fp.cell_mat[e] == 0 && fp.cell_row[e] == v && (fp.cell_par[e] >= 0 || fp.cell_val[e] != 0.0)) target = true;
        if (target) continue;
        const double m1 = b0[dw * dw + v] / n, m2 = b0[dw * dw + dw * p + v * p + v] / n;
        for (int e = 0; e < fp.n_cells; ++e) {
          const int par = fp.cell_par[e];
          if (par < 0) continue;
          if (fp.cell_mat[e] == 2 && fp.cell_col[e] == v) theta[par] = m1;
          if (fp.cell_mat[e] == 1 && fp.cell_row[e] == v && fp.cell_col[e] == v) {
            const double lb = fp.par_lbound[par];
            theta[par] = (lb == lb) ? fmax(lb, m2 - m1 * m1) : m2 - m1 * m1;
          }
        }
      }
    }
    gsync<LANES>();
  };
  load_start(fp.par_warm ? fp.par_warm : fp.par_start);
  double F = ml_eval(cx, theta, g);
  if (!isfinite(F) && fp.par_warm) {  // the warm start does not suit this SNP: the model's own start values
    load_start(fp.par_start);
    F = ml_eval(cx, theta, g);
  }
  int status = GWSEM_STATUS_ITERATION_LIMIT, it = 0;
  const double n_all = cx.n_g[0] + cx.n_g[1];
  if (!isfinite(F)) status = GWSEM_STATUS_NONZERO_GRADIENT;
  // Hessian of -2 lnL by differences of the analytic gradient (forward, relative step 1e-6; central on request)
  auto hessian = [&](bool central) {
    for (int k = 0; k < P; ++k) {
      const double t0 = theta[k];
      const double h = (central ? 1e-5 : 1e-6) * fmax(1.0, fabs(t0));
      gsync<LANES>();
      if (lane == 0) theta[k] = t0 + h;
      gsync<LANES>();
      ml_eval(cx, theta, g1);
      if (central) {
        if (lane == 0) theta[k] = t0 - h;
        gsync<LANES>();
        ml_eval(cx, theta, g2);
      }
      if (lane == 0) theta[k] = t0;
      gsync<LANES>();
      for (int a = lane; a < P; a += LANES) H[a * P + k] = central ? (g1[a] - g2[a]) / (2.0 * h) : (g1[a] - g[a]) / h;
    }
    gsync<LANES>();
    for (int e = lane; e < P * P; e += LANES) {
      const int a = e / P, b = e % P;
      if (b < a) { const double v = 0.5 * (H[a * P + b] + H[b * P + a]); Hc[e] = v; }
    }
    gsync<LANES>();
    for (int e = lane; e < P * P; e += LANES) {
      const int a = e / P, b = e % P;
      if (b < a) { H[a * P + b] = Hc[e]; H[b * P + a] = Hc[e]; }
    }
    gsync<LANES>();
  };
  const int max_iter = 100;
  double lam = 0.0, last_step = 1.0;
  for (it = 1; it <= max_iter && isfinite(F); ++it) {
    hessian(false);
    bool improved = false;
    double Fc = F;
    for (int attempt = 0; attempt < 30; ++attempt) {
      for (int e = lane; e < P * P; e += LANES) {
        const int a = e / P, b = e % P;
        double v = H[e];
        if (a == b) v += lam * fmax(fabs(H[e]), 1e-8);
        Hc[e] = v;
      }
      gsync<LANES>();
      if (!warp_cholesky<LANES>(Hc, P, lane)) { lam = lam > 0.0 ? lam * 10 : 1e-4; continue; }
      if (lane == 0) {
        for (int i = 0; i < P; ++i) {
          double v = -g[i];
          for (int k = 0; k < i; ++k) v -= Hc[i * P + k] * delta[k];
          delta[i] = v / Hc[i * P + i];
        }
        for (int i = P - 1; i >= 0; --i) {
          double v = delta[i];
          for (int k = i + 1; k < P; ++k) v -= Hc[k * P + i] * delta[k];
          delta[i] = v / Hc[i * P + i];
        }
      }
      gsync<LANES>();
      for (int k = lane; k < P; k += LANES) {
        double c = theta[k] + delta[k];
        const double lb = fp.par_lbound[k];
        if (lb == lb && c < lb) c = lb;
        cand[k] = c;
      }
      gsync<LANES>();
      Fc = ml_eval(cx, cand, g1);
      if (isfinite(Fc) && Fc <= F + 1e-12 * fabs(F)) { improved = true; break; }
      lam = lam > 0.0 ? lam * 10 : 1e-4;
    }
    if (!improved) {
      double gmax = 0.0;
      for (int k = 0; k < P; ++k) gmax = fmax(gmax, fabs(g[k]));
      status = gmax < 1e-4 * n_all ? GWSEM_STATUS_OK : GWSEM_STATUS_NONZERO_GRADIENT;
      last_step = 0.0;  // H was taken at this very point
      break;
    }
    double step = 0.0;
    for (int k = 0; k < P; ++k) step = fmax(step, fabs(cand[k] - theta[k]) / (fabs(theta[k]) + 1e-3));
    const double dF = F - Fc;
    gsync<LANES>();
    for (int k = lane; k < P; k += LANES) { theta[k] = cand[k]; g[k] = g1[k]; }
    gsync<LANES>();
    F = Fc;
    last_step = step;
    lam = lam > 1e-8 ? lam / 10 : 0.0;
    if (step < 1e-9 || dF <= 1e-13 * (1 + fabs(F))) { status = GWSEM_STATUS_OK; break; }
  }
  if (it > max_iter) it = max_iter;
  if (isfinite(F)) {
    // H from the last iteration was taken one step before the estimate; that is the same matrix to
    // O(step) relative.  Take it again at the estimate unless that step was tiny.
    if (last_step > 1e-5) hessian(false);
    if (warp_cholesky<LANES>(H, P, lane)) {
      for (int c = lane; c < P; c += LANES) {
        double* x = Hc + c;
        for (int i = 0; i < P; ++i) {
          double v = (i == c) ? 2.0 : 0.0;   // vcov = 2 H^-1
          for (int k = 0; k < i; ++k) v -= H[i * P + k] * x[k * P];
          x[i * P] = v / H[i * P + i];
        }
        for (int i = P - 1; i >= 0; --i) {
          double v = x[i * P];
          for (int k = i + 1; k < P; ++k) v -= H[k * P + i] * x[k * P];
          x[i * P] = v / H[i * P + i];
        }
      }
      gsync<LANES>();
      for (int idx = lane; idx < P * P; idx += LANES) vcov[idx] = Hc[idx];
    } else if (status == GWSEM_STATUS_OK) {
      status = GWSEM_STATUS_NOT_CONVEX;   // the Hessian is not positive definite at the solution
    }
  }
  for (int k = lane; k < P; k += LANES) est[k] = theta[k];
  if (lane == 0) {
    res.fit[snp] = isfinite(F) ? F : na;
    res.status[snp] = status;
    res.iterations[snp] = it;
  }
}

}  // namespace

void launch_fit_ml(gwsem_engine* e, const FitProgram& fp, int n_snps, const double* d_R, const int32_t* d_n,
                   const double* d_maf, const gwsem_result& d_res) {
  if (n_snps <= 0) return;
  const size_t smem = (size_t)ml_layout(fp.p, fp.nv, fp.P, 1 + fp.n_exo).total * sizeof(double);
  if (smem > 227 * 1024) fail("ML model too large for the per-SNP fit kernel (%zu bytes of scratch)", smem);
  GW_CUDA(cudaFuncSetAttribute(ml_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  e->prof_begin(kFamFit);
  ml_fit_kernel<<<n_snps, 32, smem, e->stream>>>(fp, n_snps, d_R, d_n, d_maf, d_res);
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamFit);
  e->launches++;
}

}  // namespace gwsem
