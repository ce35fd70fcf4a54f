// Body of marg_fit.cu, compiled once per group width: GW_MF_LANES threads work on one SNP (32 = one warp, used by the
// Jacobian-pattern probe; 128 = the whole CTA, used by the per-SNP fit), inside namespace GW_MF_NS.
namespace GW_MF_NS {

constexpr int LANES = GW_MF_LANES;

// model-implied statistic vector in the layout of mp.stat_*
__device__ __noinline__ void implied_stats(const FitProgram& fp, const MargProgram& mp, ModelEval<LANES>& ev, const WarpScratch& ws,
                              double* sm, const double* theta, double* sig, int lane) {
  const int nv = fp.nv;
  ev.implied(theta, sm + ws.sig);  // fills Z, C (and the covariance vech, unused here)
  double* Mv = sm + ws.T;          // T is free after implied()
  double* mu = Mv + nv;
  for (int e = lane; e < nv; e += LANES) Mv[e] = 0.0;
  gsync<LANES>();
  for (int e = lane; e < fp.n_cells; e += LANES)
    if (fp.cell_mat[e] == 2) Mv[fp.cell_col[e]] = fp.cell_par[e] >= 0 ? theta[fp.cell_par[e]] : fp.cell_val[e];
  gsync<LANES>();
  const double *Z = sm + ws.Z, *C = sm + ws.C;
  for (int i = lane; i < nv; i += LANES) {
    double v = 0.0;
    for (int k = 0; k < nv; ++k) v += Z[i * nv + k] * Mv[k];
    mu[i] = v;
  }
  gsync<LANES>();
  for (int e = lane; e < mp.n_stat; e += LANES) {
    const int kind = mp.stat_kind[e], a = mp.stat_a[e], b = mp.stat_b[e];
    const bool ord_a = mp.vcat[a] > 0;
    const double sda = ord_a ? sqrt(C[a * nv + a]) : 1.0;
    double v;
    if (kind == 0) v = mu[a];
    else if (kind == 1) {
      const int par = fp.thr_par[e];
      v = ((par >= 0 ? theta[par] : fp.thr_val[e]) - mu[a]) / sda;
    } else if (kind == 2) v = Z[a * nv + fp.exo_latent[b]] / sda;
    else if (kind == 3) v = C[a * nv + a];
    else {
      const bool ord_b = mp.vcat[b] > 0;
      const double sdb = ord_b ? sqrt(C[b * nv + b]) : 1.0;
      v = C[a * nv + b] / (sda * sdb);
    }
    sig[e] = v;
  }
  gsync<LANES>();
}

// Analytic Jacobian of the model-implied statistics (RAM derivatives chained through the
// standardisation of ordinal rows) at thv: out[i * n_c + c] = d sigma_{rows[i]} / d theta_{cols[c]}
// (rows / cols null: all of them); `pairs`: out[i] = d sigma_{rows[i]} / d theta_{cols[i]}.
__device__ __noinline__ void marg_jacobian(const FitProgram& fp, const MargProgram& mp, ModelEval<LANES>& ev, const WarpScratch& ws,
                              double* sm, const double* thv, double* sig, double* out, const int32_t* rows, int n_r,
                              const int32_t* cols, int n_c, bool pairs, int lane, bool fresh = true) {
  const int nv = fp.nv;
  if (fresh) implied_stats(fp, mp, ev, ws, sm, thv, sig, lane);   // else: the scratch already holds Z, C, mu, sig at thv
  const double *Z = sm + ws.Z, *C = sm + ws.C, *muv = sm + ws.T + nv;
  const int total = pairs ? n_r : n_r * n_c;
  for (int idx = lane; idx < total; idx += LANES) {
    const int ri = pairs ? idx : idx / n_c, ci = pairs ? idx : idx - ri * n_c;
    const int e = rows ? rows[ri] : ri, k = cols ? cols[ci] : ci;
    const int kind = mp.stat_kind[e], a = mp.stat_a[e], b = mp.stat_b[e];
    const bool ord_a = mp.vcat[a] > 0;
    const bool ord_b = kind == 4 && mp.vcat[b] > 0;
    const int xa = kind == 2 ? fp.exo_latent[b] : 0;
    double dCaa = 0.0, dCbb = 0.0, dCab = 0.0, dmu = 0.0, dZ = 0.0;
    for (int ce = fp.par_cell_ptr[k]; ce < fp.par_cell_ptr[k + 1]; ++ce) {
      const int c_id = fp.par_cells[ce];
      const int m = fp.cell_mat[c_id], rr = fp.cell_row[c_id], cc = fp.cell_col[c_id];
      if (m == 0) {
        dCaa += 2.0 * Z[a * nv + rr] * C[cc * nv + a];
        dmu += Z[a * nv + rr] * muv[cc];
        if (kind == 2) dZ += Z[a * nv + rr] * Z[cc * nv + xa];
        if (kind == 4) {
          dCbb += 2.0 * Z[b * nv + rr] * C[cc * nv + b];
          dCab += Z[a * nv + rr] * C[cc * nv + b] + Z[b * nv + rr] * C[cc * nv + a];
        }
      } else if (m == 1) {
        dCaa += Z[a * nv + rr] * Z[a * nv + cc];
        if (kind == 4) {
          dCbb += Z[b * nv + rr] * Z[b * nv + cc];
          dCab += Z[a * nv + rr] * Z[b * nv + cc];
        }
      } else {
        dmu += Z[a * nv + cc];
      }
    }
    const double Caa = C[a * nv + a];
    const double sda = ord_a ? sqrt(Caa) : 1.0;
    const double dsa = ord_a ? dCaa / (2.0 * sda * sda * sda) : 0.0;   // -d(1/sd_a)
    double v;
    if (kind == 0) v = dmu;
    else if (kind == 1) {
      const int par = fp.thr_par[e];
      const double tau = par >= 0 ? thv[par] : fp.thr_val[e];
      v = ((par == k ? 1.0 : 0.0) - dmu) / sda - (tau - muv[a]) * dsa;
    } else if (kind == 2) v = dZ / sda - Z[a * nv + xa] * dsa;
    else if (kind == 3) v = dCaa;
    else {
      const double Cbb = C[b * nv + b], sdb = ord_b ? sqrt(Cbb) : 1.0;
      const double dsb = ord_b ? dCbb / (2.0 * sdb * sdb * sdb) : 0.0;
      v = dCab / (sda * sdb) - C[a * nv + b] * (dsa / sdb + dsb / sda);
    }
    out[idx] = v;
  }
  gsync<LANES>();
}

// In-place inverse from the Cholesky factor: H (lower factor, n x n) -> out = (L L')^-1, column per lane
__device__ void chol_inverse(const double* H, int n, double* out, int lane) {
  for (int c = lane; c < n; c += LANES) {
    double* x = out + c;
    for (int i = 0; i < n; ++i) {
      double v = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) v -= H[i * n + k] * x[k * n];
      x[i * n] = v / H[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double v = x[i * n];
      for (int k = i + 1; k < n; ++k) v -= H[k * n + i] * x[k * n];
      x[i * n] = v / H[i * n + i];
    }
  }
  gsync<LANES>();
}

__global__ void __launch_bounds__(GW_MF_LANES > 64 ? GW_MF_LANES : 64)
    marg_fit_kernel(FitProgram fp, MargProgram mp, int n_snps, const double* __restrict__ summary,
                    const double* __restrict__ maf_in, gwsem_result res) {
  extern __shared__ double smem_d[];
  const int group = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int snp = blockIdx.x * (blockDim.x / LANES) + group;
  if (snp >= n_snps) return;
  const int J = mp.J, n1 = mp.n1, q0 = mp.q0, D = 2 + J;
  const int nt = mp.nt, ns = mp.n_stat, P = fp.P, nv = fp.nv;
  const WarpScratch ws = marg_ws(fp.p, nv);
  const MargScratch ms = marg_layout(nt, ns, P, mp.n_a > 0 ? mp.n_b : 0);
  double* sm = smem_d + (size_t)group * (ws.total + ms.total);
  double* mx = sm + ws.total;
  const double* rec = summary + (int64_t)snp * mp.sum_stride;
  // full record (marg_exact.cu): every estimate and score product was taken over this SNP's retained people
  const double* frec = mp.full ? mp.full + (int64_t)snp * mp.full_stride : nullptr;
  const bool full = frec && (mp.full_all || frec[0] >= 0.0);
  if (mp.fast_fit && !full) return;   // marg_fit_fast_kernel has this SNP
  const int n = (int)(full ? frec[0] : rec[0]);
  const double mu_g = full ? 0.0 : rec[1], var_g = full ? 0.0 : rec[2];

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
  if (full) {
    if (frec[1] != 0.0) {
      if (lane == 0) { res.catch_code[snp] = (int)frec[1]; res.catch_var[snp] = (int)frec[2]; }
      return;
    }
  } else if (n < 2 || !(var_g * (double)n / (double)(n - 1) >= fp.min_variance)) {
    if (lane == 0) { res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE; res.catch_var[snp] = 0; }
    return;
  }

  // ---- theta, A, B.   index map: [0,1] snp | 2.. items | ir = 2+n1 .. rho(snp,item) | ii .. rho(item,item)
  const int ir = 2 + n1, ii = ir + J;
  double *th = mx + ms.th, *A = mx + ms.A, *B = mx + ms.B, *Ai = mx + ms.Ai, *T = mx + ms.T;
  if (full) {
    const int n1t = mp.n1t;
    for (int e = lane; e < nt; e += LANES) th[e] = frec[mp.fo_theta + e];
    for (int e = lane; e < nt * nt; e += LANES) {
      const int x = e / nt, y = e % nt;
      const double b = frec[mp.fo_B + e];
      B[e] = b;
      // A: block-diagonal stage-1 information (outer products), A21 rows, diagonal A22
      double a = 0.0;
      if (x < n1t && y < n1t) {
        int vx = 0, vy = 0;
        while (x >= mp.voff[vx + 1]) ++vx;
        while (y >= mp.voff[vy + 1]) ++vy;
        if (vx == vy) a = b;
      } else if (x >= n1t) {
        if (y < n1t) a = frec[mp.fo_a21 + (x - n1t) * n1t + y];
        else if (y == x) a = b;
      }
      A[e] = a;
      Ai[e] = x == y ? 1.0 : 0.0;
    }
  } else {
  for (int e = lane; e < nt; e += LANES) {
    double v;
    if (e == 0) v = mu_g;
    else if (e == 1) v = var_g;
    else if (e < ir) v = mp.theta1[e - 2];
    else if (e < ii) v = rec[mp.o_rho + (e - ir)];
    else v = mp.rho_ii[e - ii];
    th[e] = v;
  }
  // column x of theta as a score column: SNP-dependent set D (0,1 -> 0,1; ir+j -> 2+j) or SC0 column
  auto dcol = [&](int x) { return x < 2 ? x : (x >= ir && x < ii ? 2 + (x - ir) : -1); };
  auto ccol = [&](int x) { return x >= 2 && x < ir ? x - 2 : (x >= ii ? n1 + (x - ii) : -1); };
  for (int e = lane; e < nt * nt; e += LANES) {
    const int x = e / nt, y = e % nt;
    const int dx = dcol(x), dy = dcol(y), cx = ccol(x), cy = ccol(y);
    double v;
    if (dx >= 0 && dy >= 0) v = rec[mp.o_bdd + dx * D + dy];
    else if (dx >= 0) v = rec[mp.o_bd0 + dx * q0 + cy];
    else if (dy >= 0) v = rec[mp.o_bd0 + dy * q0 + cx];
    else v = mp.B00[cx * q0 + cy];
    B[e] = v;
    // A: block diagonal stage-1 information (outer products), A21 rows, diagonal A22
    double a = 0.0;
    if (x < 2 && y < 2) a = rec[mp.o_bdd + x * D + y];
    else if (x >= 2 && x < ir && y >= 2 && y < ir) {
      int jx = -1, jy = -1;
      for (int j = 0; j < J; ++j) {
        if (x - 2 >= mp.col_of[j] && x - 2 < mp.col_of[j + 1]) jx = j;
        if (y - 2 >= mp.col_of[j] && y - 2 < mp.col_of[j + 1]) jy = j;
      }
      if (jx == jy) a = mp.B00[(x - 2) * q0 + (y - 2)];
    } else if (x >= ir && x < ii) {
      const int j = x - ir;
      if (y < 2) a = rec[mp.o_a21snp + 2 * j + y];
      else if (y < ir && y - 2 >= mp.col_of[j] && y - 2 < mp.col_of[j + 1]) a = rec[mp.o_a21item + (y - 2)];
      else if (y == x) a = rec[mp.o_a22 + j];
    } else if (x >= ii) {
      if (y >= 2 && y < ir) a = mp.A21ii[(x - ii) * n1 + (y - 2)];
      else if (y == x) a = mp.B00[(n1 + x - ii) * q0 + (n1 + x - ii)];
    }
    A[e] = a;
    Ai[e] = x == y ? 1.0 : 0.0;
  }
  }
  gsync<LANES>();

  // ---- Ai = A^-1 by Gauss-Jordan with partial pivoting (A is block lower triangular, well conditioned)
  bool singular = false;
  for (int c = 0; c < nt; ++c) {
    int piv = c;
    double best = fabs(A[c * nt + c]);
    for (int r = c + 1; r < nt; ++r)
      if (fabs(A[r * nt + c]) > best) { best = fabs(A[r * nt + c]); piv = r; }
    if (!(best > 0.0)) { singular = true; break; }
    gsync<LANES>();   // everybody has read column c (and agrees on the pivot) before the swap rewrites it
    if (piv != c) {
      for (int k = lane; k < nt; k += LANES) {
        double t = A[c * nt + k]; A[c * nt + k] = A[piv * nt + k]; A[piv * nt + k] = t;
        t = Ai[c * nt + k]; Ai[c * nt + k] = Ai[piv * nt + k]; Ai[piv * nt + k] = t;
      }
      gsync<LANES>();
    }
    const double inv = 1.0 / A[c * nt + c];
    gsync<LANES>();
    for (int k = lane; k < nt; k += LANES) { A[c * nt + k] *= inv; Ai[c * nt + k] *= inv; }
    gsync<LANES>();
    // eliminate column c from every other row in one sweep: the multipliers A[r][c] go to a side buffer first
    // (T is idle until the inverse is done), so nobody reads a multiplier that somebody else has already cleared
    for (int r = lane; r < nt; r += LANES) T[r] = r == c ? 0.0 : A[r * nt + c];
    gsync<LANES>();
    for (int e = lane; e < nt * nt; e += LANES) {
      const double f = T[e / nt];
      if (f == 0.0) continue;
      const int k = e % nt;
      A[e] -= f * A[c * nt + k];
      Ai[e] -= f * Ai[c * nt + k];
    }
    gsync<LANES>();
  }
  if (singular) {
    if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
    return;
  }
  // T = Ai * B ; A <- acov = T * Ai'
  for (int e = lane; e < nt * nt; e += LANES) {
    const int x = e / nt, y = e % nt;
    double v = 0.0;
    for (int k = 0; k < nt; ++k) v += Ai[x * nt + k] * B[k * nt + y];
    T[e] = v;
  }
  gsync<LANES>();
  double* acov = A;
  for (int e = lane; e < nt * nt; e += LANES) {
    const int x = e / nt, y = e % nt;
    double v = 0.0;
    // lanes hold different rows y of Ai (a stride of nt doubles: one bank when nt = 32): each lane starts its walk
    // along k at its own y, so that a step touches nt different banks
    for (int k = 0; k < nt; ++k) {
      int kk = k + y;
      if (kk >= nt) kk -= nt;
      v += T[x * nt + kk] * Ai[y * nt + kk];
    }
    acov[e] = v;
  }
  gsync<LANES>();

  // ---- statistics and their delta-method rows (<= 3 non-zeros each)
  double *s = mx + ms.s, *hidx = mx + ms.hidx, *hcoef = mx + ms.hcoef, *L = mx + ms.L, *gscale = mx + ms.scale;
  // sd of a continuous manifest variable: its stage-1 variance is the last entry of its block of theta
  auto sd_of = [&](int a) { return sqrt(th[mp.voff[a + 1] - 1]); };
  for (int e = lane; e < ns; e += LANES) {
    const int kind = mp.stat_kind[e], x = mp.stat_theta[e];
    double i0 = x, c0 = 1.0, i1 = 0, c1 = 0.0, i2 = 0, c2 = 0.0, v = th[x];
    if (kind == 4) {  // covariance of variables a > b: rho * sd_a * sd_b, sd = 1 for ordinal
      const int a = mp.stat_a[e], b = mp.stat_b[e];
      const bool ca = mp.vcat[a] == 0, cb = mp.vcat[b] == 0;
      const double sa = ca ? sd_of(a) : 1.0, sb = cb ? sd_of(b) : 1.0;
      const double rho = th[x];
      v = rho * sa * sb;
      c0 = sa * sb;
      // variance parameter of a continuous variable: snp -> theta 1; item j -> last column of its block
      if (ca) { i1 = mp.voff[a + 1] - 1; c1 = rho * sb / (2.0 * sa); }
      if (cb) { i2 = mp.voff[b + 1] - 1; c2 = rho * sa / (2.0 * sb); }
    }
    s[e] = v;
    hidx[3 * e] = i0; hidx[3 * e + 1] = i1; hidx[3 * e + 2] = i2;
    hcoef[3 * e] = c0; hcoef[3 * e + 1] = c1; hcoef[3 * e + 2] = c2;
  }
  gsync<LANES>();
  for (int e = lane; e < ns * ns; e += LANES) {
    const int x = e / ns, y = e % ns;
    double v = 0.0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        v += hcoef[3 * x + a] * hcoef[3 * y + b] * acov[(int)hidx[3 * x + a] * nt + (int)hidx[3 * y + b]];
    L[e] = v;
  }
  gsync<LANES>();
  for (int e = lane; e < ns; e += LANES) gscale[e] = fabs(L[e * ns + e]);
  gsync<LANES>();
  for (int e = lane; e < ns * ns; e += LANES) {  // symmetrise before factorising
    const int x = e / ns, y = e % ns;
    if (y < x) { const double v = 0.5 * (L[x * ns + y] + L[y * ns + x]); L[x * ns + y] = v; }
  }
  gsync<LANES>();
  for (int e = lane; e < ns * ns; e += LANES) {
    const int x = e / ns, y = e % ns;
    if (y > x) L[e] = L[y * ns + x];
  }
  gsync<LANES>();
  // ---- Fit.  Parameters that enter exactly one statistic (thresholds, covariate slopes: mp.elim_a_*)
  // can reproduce any value of it, so minimising over them leaves the b-statistics' residuals
  // weighted by the inverse of their own covariance block (a Schur complement of the full weight):
  //   min_theta r' Gamma^-1 r  =  min_beta r_b' Gamma_bb^-1 r_b ,   r_a = Gamma_ab Gamma_bb^-1 r_b .
  // The iteration runs on beta only (n_beta parameters, n_b statistics); the eliminated parameters
  // follow in closed form and the full Jacobian is needed once, for the parameter covariance.
  const int n_el = mp.n_a, nb = n_el > 0 ? mp.n_b : ns, nbeta = n_el > 0 ? mp.n_beta : P;
  const int32_t* b_row = n_el > 0 ? mp.elim_b_row : nullptr;
  const int32_t* b_par = n_el > 0 ? mp.elim_b_par : nullptr;
  double* Lb = n_el > 0 ? mx + ms.Lb : L;
  if (n_el > 0) {
    for (int e = lane; e < nb * nb; e += LANES) Lb[e] = L[b_row[e / nb] * ns + b_row[e % nb]];
    double* sc = mx + ms.sigp;  // pivot scales of the b block
    for (int e = lane; e < nb; e += LANES) sc[e] = gscale[b_row[e]];
    gsync<LANES>();
    if (!warp_cholesky<LANES>(Lb, nb, lane, sc)) {
      if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
      return;
    }
  } else if (!warp_cholesky<LANES>(L, ns, lane, gscale)) {
    if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
    return;
  }

  // ---- damped Gauss-Newton on beta (same schedule as fit.cu / the oracle)
  double *theta = mx + ms.theta, *cand = mx + ms.cand, *sig = mx + ms.sig, *r = mx + ms.r;
  double *rc = mx + ms.rc, *Jm = mx + ms.J, *g = mx + ms.g, *H = mx + ms.H, *Hc = mx + ms.Hc, *delta = mx + ms.delta;
  ModelEval<LANES> ev{fp, ws, sm, lane};
  auto objective = [&](const double* thv, double* resid) -> double {
    implied_stats(fp, mp, ev, ws, sm, thv, sig, lane);
    for (int e = lane; e < nb; e += LANES) { const int row = b_row ? b_row[e] : e; resid[e] = s[row] - sig[row]; }
    gsync<LANES>();
    warp_forward_solve<LANES>(Lb, nb, resid, 1, 1, lane);
    double f = 0.0;
    for (int e = 0; e < nb; ++e) f += resid[e] * resid[e];
    return f;
  };
  for (int k = lane; k < P; k += LANES) theta[k] = fp.par_warm ? fp.par_warm[k] : fp.par_start[k];
  gsync<LANES>();
  double F = objective(theta, r);
  double lam = 1e-4;
  int status = GWSEM_STATUS_ITERATION_LIMIT;
  int it = 0;
  const int max_iter = 200;
  for (it = 1; it <= max_iter && nbeta > 0; ++it) {
    marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Jm, b_row, nb, b_par, nbeta, false, lane);
    warp_forward_solve<LANES>(Lb, nb, Jm, nbeta, nbeta, lane);
    for (int k = lane; k < nbeta; k += LANES) {
      double v = 0.0;
      for (int e = 0; e < nb; ++e) v += Jm[e * nbeta + k] * r[e];
      g[k] = v;
    }
    for (int idx = lane; idx < nbeta * nbeta; idx += LANES) {
      const int a = idx / nbeta, b = idx % nbeta;
      double v = 0.0;
      for (int e = 0; e < nb; ++e) v += Jm[e * nbeta + a] * Jm[e * nbeta + b];
      H[idx] = v;
    }
    gsync<LANES>();
    bool improved = false;
    double Fc = F;
    for (int attempt = 0; attempt < 30; ++attempt) {
      for (int idx = lane; idx < nbeta * nbeta; idx += LANES) {
        const int a = idx / nbeta, b = idx % nbeta;
        double v = H[idx];
        if (a == b) v += lam * fmax(H[idx], 1e-12);
        Hc[idx] = v;
      }
      gsync<LANES>();
      if (!warp_cholesky<LANES>(Hc, nbeta, lane)) { lam *= 10; continue; }
      if (lane == 0) {
        for (int i = 0; i < nbeta; ++i) {
          double v = g[i];
          for (int k = 0; k < i; ++k) v -= Hc[i * nbeta + k] * delta[k];
          delta[i] = v / Hc[i * nbeta + i];
        }
        for (int i = nbeta - 1; i >= 0; --i) {
          double v = delta[i];
          for (int k = i + 1; k < nbeta; ++k) v -= Hc[k * nbeta + i] * delta[k];
          delta[i] = v / Hc[i * nbeta + i];
        }
      }
      gsync<LANES>();
      for (int k = lane; k < P; k += LANES) cand[k] = theta[k];
      gsync<LANES>();
      for (int c = lane; c < nbeta; c += LANES) {
        const int k = b_par ? b_par[c] : c;
        double v = theta[k] + delta[c];
        const double lb = fp.par_lbound[k];
        if (lb == lb && v < lb) v = lb;
        cand[k] = v;
      }
      gsync<LANES>();
      Fc = objective(cand, rc);
      if (isfinite(Fc) && Fc <= F) { improved = true; break; }
      lam *= 10;
    }
    if (!improved) {
      double gmax = 0.0;
      for (int k = 0; k < nbeta; ++k) gmax = fmax(gmax, fabs(g[k]));
      status = gmax < 1e-6 * (1 + F) ? GWSEM_STATUS_OK : GWSEM_STATUS_NONZERO_GRADIENT;
      break;
    }
    double step = 0.0;
    for (int c = 0; c < nbeta; ++c) {
      const int k = b_par ? b_par[c] : c;
      step = fmax(step, fabs(cand[k] - theta[k]) / (fabs(theta[k]) + 1e-3));
    }
    const double dF = F - Fc;
    gsync<LANES>();
    for (int k = lane; k < P; k += LANES) theta[k] = cand[k];
    for (int e = lane; e < nb; e += LANES) r[e] = rc[e];
    gsync<LANES>();
    F = Fc;
    lam = fmax(lam / 10, 1e-12);
    if (step < 1e-10 || dF <= 1e-14 * (1 + F)) { status = GWSEM_STATUS_OK; break; }
  }
  if (nbeta == 0) { status = GWSEM_STATUS_OK; it = 0; }
  if (it > max_iter) it = max_iter;

  if (n_el > 0) {
    // eliminated parameters: r_a = Gamma_ab Gamma_bb^-1 r_b, then each one from its own statistic
    // (r holds Lb^-1 r_b; back-substitution gives w = Gamma_bb^-1 r_b)
    double* w = rc;
    if (lane == 0)
      for (int i = nb - 1; i >= 0; --i) {
        double v = r[i];
        for (int k = i + 1; k < nb; ++k) v -= Lb[k * nb + i] * w[k];
        w[i] = v / Lb[i * nb + i];
      }
    gsync<LANES>();
    double* target = mx + ms.sigp;  // n_el <= ns
    for (int i = lane; i < n_el; i += LANES) {
      const int row = mp.elim_a_row[i];
      double ra = 0.0;
      for (int j = 0; j < nb; ++j) ra += L[row * ns + b_row[j]] * w[j];   // L still holds Gamma
      target[i] = s[row] - ra;
    }
    gsync<LANES>();
    for (int round = 0; round < 2; ++round) {  // exact after one round when the statistic is linear in its parameter
      marg_jacobian(fp, mp, ev, ws, sm, theta, sig, delta, mp.elim_a_row, n_el, mp.elim_a_par, n_el, true, lane);
      for (int i = lane; i < n_el; i += LANES) {
        const double d = delta[i];
        if (d != 0.0) theta[mp.elim_a_par[i]] += (target[i] - sig[mp.elim_a_row[i]]) / d;
      }
      gsync<LANES>();
    }
    if (!warp_cholesky<LANES>(L, ns, lane, gscale)) {
      if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
      return;
    }
  }

  // ---- parameter covariance (J' W J)^-1 from the full Jacobian at the estimate
  marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Jm, nullptr, ns, nullptr, P, false, lane);
  warp_forward_solve<LANES>(L, ns, Jm, P, P, lane);
  for (int idx = lane; idx < P * P; idx += LANES) {
    const int a = idx / P, b = idx % P;
    double v = 0.0;
    for (int e = 0; e < ns; ++e) v += Jm[e * P + a] * Jm[e * P + b];
    H[idx] = v;
  }
  gsync<LANES>();
  if (warp_cholesky<LANES>(H, P, lane)) {
    chol_inverse(H, P, Hc, lane);
    for (int idx = lane; idx < P * P; idx += LANES) vcov[idx] = Hc[idx];
  } else if (status == GWSEM_STATUS_OK) {
    status = GWSEM_STATUS_NOT_CONVEX;   // the information matrix is not positive definite at the solution
  }
  for (int k = lane; k < P; k += LANES) est[k] = theta[k];
  if (lane == 0) {
    res.fit[snp] = F;
    res.status[snp] = status;
    res.iterations[snp] = it;
  }
}

#if GW_MF_LANES == 32
// ---- The structured fit: one warp per SNP, for records assembled from the SNP-independent item side (marg_series /
// marg_stats) of a model with eliminated parameters.  Same estimator as marg_fit_kernel, with the algebra that does
// not depend on the SNP taken out of the loop:
//  * theta splits into the D = 2 + J SNP-side entries d (mean, variance, rho(snp, item_j)) and the q0 item-side entries
//    c.  A is block lower triangular with A_cd = 0, so (A^-1)_cc = W = A_cc^-1 and acov_cc = W B_cc W' are constants
//    (mp.Wcc, mp.acov_cc: once per model on the host); the d rows of A^-1 are A_dd^-1 [I, -Q] with Q = A_dc W, and
//    A_dc only links rho(snp, item_j) to item j's own stage-1 block.  Per SNP: acov_d. = A_dd^-1 ([I, -Q] B) A^-T,
//    D rows instead of nt^3 work.
//  * with the a-statistics (one eliminated parameter each) and the b-statistics, X = Gamma_ab Lb^-T (Lb Lb' =
//    Gamma_bb), Jt = Lb^-1 J_b,beta, K = J_a,beta - X Jt, V = (Jt' Jt)^-1 and Da = diag(d sigma_a / d alpha):
//      vcov(beta) = V,  cov(alpha, beta) = -Da^-1 K V,  vcov(alpha) = Da^-1 (Gamma_aa - X X' + K V K') Da^-1,
//    which is (J' Gamma^-1 J)^-1 without the ns x ns factorisation or the P x P inverse.
__device__ unsigned long long g_fast_clk[8];   // development aid (GWSEM_FIT_DIAG): cycles per phase of marg_fit_fast_kernel
struct FastScratch {
  int th, Adi, Q, NBd, NBc, Rdc, rows, s, hidx, hcoef, Lb, sc, X, theta, cand, sig, r, rc, Jb, g, H, Hc, delta, Ja, KV, Da, S, total;
};
__host__ __device__ inline FastScratch fast_layout(int nt, int D, int q0, int ns, int P, int na, int nb, int nbeta) {
  FastScratch m;
  int o = 0;
  auto take = [&](int n) { int at = o; o += (n + 1) & ~1; return at; };
  m.th = take(nt);
  // {Adi, Q, NBd, NBc, Rdc} are dead once `rows` is complete; the Schur complement S takes their place
  {
    const int start = o;
    m.Adi = take(D * D); m.Q = take(D * q0); m.NBd = take(D * D); m.NBc = take(D * q0); m.Rdc = take(D * q0);
    m.S = start;
    if (o - start < na * na) o = start + na * na;
  }
  m.rows = take(D * (D + q0));
  m.s = take(ns);
  // the delta-method tables are dead once Gamma's blocks are taken; J_a,beta and K V take their place
  {
    const int start = o;
    m.hidx = take(3 * ns); m.hcoef = take(3 * ns);
    m.Ja = start; m.KV = start + na * nbeta;
    if (o - start < 2 * na * nbeta) o = start + 2 * na * nbeta;
  }
  m.Lb = take(nb * nb); m.sc = take(nb); m.X = take(na * nb);
  m.theta = take(P); m.cand = take(P); m.sig = take(ns); m.r = take(nb); m.rc = take(nb);
  m.Jb = take(nb * nbeta); m.g = take(nbeta); m.H = take(nbeta * nbeta); m.Hc = take(nbeta * nbeta); m.delta = take(na > nbeta ? na : nbeta);
  m.Da = take(na);
  m.total = o;
  return m;
}

__device__ void fast_fit_one(const FitProgram& fp, const MargProgram& mp, const WarpScratch& ws, const FastScratch& ms,
                             int snp, double* sm, const double* __restrict__ summary, const double* __restrict__ maf_in,
                             const gwsem_result& res, int lane) {
  const int J = mp.J, n1 = mp.n1, q0 = mp.q0, D = 2 + J;
  const int nt = mp.nt, ns = mp.n_stat, P = fp.P;
  const int na = mp.n_a, nb = mp.n_b, nbeta = mp.n_beta;
  double* mx = sm + ws.total;
  const double* rec = summary + (int64_t)snp * mp.sum_stride;
  const double* frec = mp.full ? mp.full + (int64_t)snp * mp.full_stride : nullptr;
  if (frec && (mp.full_all || frec[0] >= 0.0)) return;   // full record: marg_fit_kernel's
  long long clk_last = clock64();
  auto lap = [&](int slot) {
    if (!mp.diag) return;
    const long long c = clock64();
    if (lane == 0) atomicAdd(&g_fast_clk[slot], (unsigned long long)(c - clk_last));
    clk_last = c;
  };
  const int n = (int)rec[0];
  const double mu_g = rec[1], var_g = rec[2];
  double* est = res.est + (int64_t)snp * P;
  double* vcov = res.vcov + (int64_t)snp * P * P;
  const double na_val = __longlong_as_double((long long)GWSEM_NA_BITS);
  for (int k = lane; k < P; k += 32) est[k] = fp.par_start[k];
  for (int k = lane; k < P * P; k += 32) vcov[k] = na_val;
  if (lane == 0) {
    res.maf[snp] = maf_in[snp];
    res.n_used[snp] = n;
    res.fit[snp] = na_val;
    res.status[snp] = GWSEM_STATUS_NA;
    res.catch_code[snp] = GWSEM_CATCH_NONE;
    res.catch_var[snp] = -1;
    res.iterations[snp] = 0;
  }
  if (n < 2 || !(var_g * (double)n / (double)(n - 1) >= fp.min_variance)) {
    if (lane == 0) { res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE; res.catch_var[snp] = 0; }
    return;
  }
  auto not_pd = [&]() { if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD; };

  const int ir = 2 + n1, ii = ir + J;
  double *th = mx + ms.th, *Adi = mx + ms.Adi, *Q = mx + ms.Q, *NBd = mx + ms.NBd, *NBc = mx + ms.NBc, *Rdc = mx + ms.Rdc;
  double* rows = mx + ms.rows;   // acov of the D SNP-side entries against [d (D) | c (q0)]
  const int RW = D + q0;
  for (int e = lane; e < nt; e += 32) {
    double v;
    if (e == 0) v = mu_g;
    else if (e == 1) v = var_g;
    else if (e < ir) v = mp.theta1[e - 2];
    else if (e < ii) v = rec[mp.o_rho + (e - ir)];
    else v = mp.rho_ii[e - ii];
    th[e] = v;
  }
  // A_dd^-1: [[S, 0], [R, diag(a22)]]^-1 = [[S^-1, 0], [-diag(1/a22) R S^-1, diag(1/a22)]]
  const double* bdd = rec + mp.o_bdd;
  const double s00 = bdd[0], s01 = bdd[1], s10 = bdd[D], s11 = bdd[D + 1], det = s00 * s11 - s01 * s10;
  bool singular = !(det != 0.0) || !isfinite(det);
  for (int j = 0; j < J; ++j) singular = singular || !(rec[mp.o_a22 + j] != 0.0);
  if (singular) { not_pd(); return; }
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    const double i00 = s11 / det, i01 = -s01 / det, i10 = -s10 / det, i11 = s00 / det;
    double v = 0.0;
    if (x < 2 && y < 2) v = x == 0 ? (y == 0 ? i00 : i01) : (y == 0 ? i10 : i11);
    else if (x >= 2 && y < 2) {
      const int j = x - 2;
      const double r0 = rec[mp.o_a21snp + 2 * j], r1 = rec[mp.o_a21snp + 2 * j + 1];
      v = -(r0 * (y == 0 ? i00 : i01) + r1 * (y == 0 ? i10 : i11)) / rec[mp.o_a22 + j];
    } else if (x == y) v = 1.0 / rec[mp.o_a22 + (x - 2)];
    Adi[e] = v;
  }
  // Q = A_dc W: row 2 + j lives on item j's own stage-1 block
  for (int e = lane; e < D * q0; e += 32) {
    const int x = e / q0, k = e % q0;
    double v = 0.0;
    if (x >= 2) {
      const int j = x - 2, c0 = mp.col_of[j], c1 = mp.col_of[j + 1];
      if (k >= c0 && k < c1)
        for (int m = c0; m < c1; ++m) v += rec[mp.o_a21item + m] * mp.Wcc[m * q0 + k];
    }
    Q[e] = v;
  }
  __syncwarp();
  // NB = [I, -Q] B, split into its d and c columns
  const double* bd0 = rec + mp.o_bd0;
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    double v = bdd[x * D + y];
    if (x >= 2) {
      const int j = x - 2;
      for (int k = mp.col_of[j]; k < mp.col_of[j + 1]; ++k) v -= Q[x * q0 + k] * bd0[y * q0 + k];
    }
    NBd[e] = v;
  }
  for (int e = lane; e < D * q0; e += 32) {
    const int x = e / q0, y = e % q0;
    double v = bd0[x * q0 + y];
    if (x >= 2) {
      const int j = x - 2;
      for (int k = mp.col_of[j]; k < mp.col_of[j + 1]; ++k) v -= Q[x * q0 + k] * mp.B00[k * q0 + y];
    }
    NBc[e] = v;
  }
  __syncwarp();
  // R_dc = NB_c W' ; R_dd = NB_d - NB_c Q' (into the d columns of `rows`, in place of the final values for now)
  for (int e = lane; e < D * q0; e += 32) {
    const int x = e / q0, y = e % q0;
    double v = 0.0;
    for (int k = 0; k < q0; ++k) v += NBc[x * q0 + k] * mp.Wcc[y * q0 + k];
    Rdc[e] = v;
  }
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    double v = NBd[e];
    if (y >= 2) {
      const int j = y - 2;
      for (int k = mp.col_of[j]; k < mp.col_of[j + 1]; ++k) v -= NBc[x * q0 + k] * Q[y * q0 + k];
    }
    NBd[e] = v;   // each entry read and written by the same lane
  }
  __syncwarp();
  // acov_dc = A_dd^-1 R_dc ; acov_dd = A_dd^-1 R_dd A_dd^-T (symmetrised)
  for (int e = lane; e < D * q0; e += 32) {
    const int x = e / q0, y = e % q0;
    double v = 0.0;
    for (int k = 0; k < D; ++k) v += Adi[x * D + k] * Rdc[k * q0 + y];
    rows[x * RW + D + y] = v;
  }
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    double v = 0.0;
    for (int k = 0; k < D; ++k)
      for (int m = 0; m < D; ++m) v += Adi[x * D + k] * NBd[k * D + m] * Adi[y * D + m];
    rows[x * RW + y] = v;
  }
  __syncwarp();
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    if (y < x) { const double v = 0.5 * (rows[x * RW + y] + rows[y * RW + x]); rows[x * RW + y] = v; }
  }
  __syncwarp();
  for (int e = lane; e < D * D; e += 32) {
    const int x = e / D, y = e % D;
    if (y > x) rows[x * RW + y] = rows[y * RW + x];
  }
  __syncwarp();
  auto acov_at = [&](int x, int y) -> double {
    const int dx = x < 2 ? x : (x >= ir && x < ii ? 2 + (x - ir) : -1), dy = y < 2 ? y : (y >= ir && y < ii ? 2 + (y - ir) : -1);
    const int cx = x >= 2 && x < ir ? x - 2 : (x >= ii ? n1 + (x - ii) : -1), cy = y >= 2 && y < ir ? y - 2 : (y >= ii ? n1 + (y - ii) : -1);
    if (dx >= 0) return dy >= 0 ? rows[dx * RW + dy] : rows[dx * RW + D + cy];
    if (dy >= 0) return rows[dy * RW + D + cx];
    return mp.acov_cc[cx * q0 + cy];
  };

  lap(0);
  // ---- statistics and their delta-method rows (as in marg_fit_kernel)
  double *s = mx + ms.s, *hidx = mx + ms.hidx, *hcoef = mx + ms.hcoef;
  auto sd_of = [&](int a) { return sqrt(th[mp.voff[a + 1] - 1]); };
  for (int e = lane; e < ns; e += 32) {
    const int kind = mp.stat_kind[e], x = mp.stat_theta[e];
    double i0 = x, c0 = 1.0, i1 = 0, c1 = 0.0, i2 = 0, c2 = 0.0, v = th[x];
    if (kind == 4) {
      const int a = mp.stat_a[e], b = mp.stat_b[e];
      const bool ca = mp.vcat[a] == 0, cb = mp.vcat[b] == 0;
      const double sa = ca ? sd_of(a) : 1.0, sb = cb ? sd_of(b) : 1.0;
      const double rho = th[x];
      v = rho * sa * sb;
      c0 = sa * sb;
      if (ca) { i1 = mp.voff[a + 1] - 1; c1 = rho * sb / (2.0 * sa); }
      if (cb) { i2 = mp.voff[b + 1] - 1; c2 = rho * sa / (2.0 * sb); }
    }
    s[e] = v;
    hidx[3 * e] = i0; hidx[3 * e + 1] = i1; hidx[3 * e + 2] = i2;
    hcoef[3 * e] = c0; hcoef[3 * e + 1] = c1; hcoef[3 * e + 2] = c2;
  }
  __syncwarp();
  auto gamma_at = [&](int x, int y) -> double {   // covariance of statistics x and y
    double v = 0.0;
    for (int a = 0; a < 3; ++a) {
      const double ca = hcoef[3 * x + a];
      if (ca == 0.0) continue;
      for (int b = 0; b < 3; ++b) {
        const double cb = hcoef[3 * y + b];
        if (cb != 0.0) v += ca * cb * acov_at((int)hidx[3 * x + a], (int)hidx[3 * y + b]);
      }
    }
    return v;
  };
  const int32_t *b_row = mp.elim_b_row, *b_par = mp.elim_b_par, *a_row = mp.elim_a_row, *a_par = mp.elim_a_par;
  double *Lb = mx + ms.Lb, *sc = mx + ms.sc, *X = mx + ms.X;
  for (int e = lane; e < nb * nb; e += 32) {
    const int x = e / nb, y = e % nb;
    if (y <= x) { const double v = gamma_at(b_row[x], b_row[y]); Lb[x * nb + y] = v; Lb[y * nb + x] = v; }
  }
  for (int e = lane; e < na * nb; e += 32) X[e] = gamma_at(a_row[e / nb], b_row[e % nb]);
  __syncwarp();
  for (int e = lane; e < nb; e += 32) sc[e] = fabs(Lb[e * nb + e]);
  __syncwarp();
  if (!warp_cholesky<32>(Lb, nb, lane, sc)) { not_pd(); return; }
  // X <- Gamma_ab Lb^-T (row i: forward substitution against Lb)
  for (int i = lane; i < na; i += 32) {
    double* x = X + i * nb;
    for (int c = 0; c < nb; ++c) {
      double v = x[c];
      for (int k = 0; k < c; ++k) v -= Lb[c * nb + k] * x[k];
      x[c] = v / Lb[c * nb + c];
    }
  }
  __syncwarp();
  // Schur complement Gamma_aa - X X' (kept for vcov(alpha); its factorisation below is the positive-definiteness check
  // of the whole Gamma that marg_fit_kernel makes)
  double* S = mx + ms.S;
  for (int e = lane; e < na * na; e += 32) {
    const int x = e / na, y = e % na;
    if (y > x) continue;
    double v = gamma_at(a_row[x], a_row[y]);
    for (int k = 0; k < nb; ++k) v -= X[x * nb + k] * X[y * nb + k];
    S[x * na + y] = v;
    S[y * na + x] = v;
  }
  __syncwarp();

  lap(1);
  // ---- damped Gauss-Newton on beta (same schedule as marg_fit_kernel)
  double *theta = mx + ms.theta, *cand = mx + ms.cand, *sig = mx + ms.sig, *r = mx + ms.r, *rc = mx + ms.rc;
  double *Jm = mx + ms.Jb, *g = mx + ms.g, *H = mx + ms.H, *Hc = mx + ms.Hc, *delta = mx + ms.delta;
  ModelEval<32> ev{fp, ws, sm, lane};
  auto objective = [&](const double* thv, double* resid) -> double {
    implied_stats(fp, mp, ev, ws, sm, thv, sig, lane);
    for (int e = lane; e < nb; e += 32) resid[e] = s[b_row[e]] - sig[b_row[e]];
    __syncwarp();
    warp_forward_solve<32>(Lb, nb, resid, 1, 1, lane);
    double f = 0.0;
    for (int e = 0; e < nb; ++e) f += resid[e] * resid[e];
    return f;
  };
  for (int k = lane; k < P; k += 32) theta[k] = fp.par_warm ? fp.par_warm[k] : fp.par_start[k];
  __syncwarp();
  if (mp.n_init > 0) {
    // parameters that are a mean / variance outright start at the observed value (the snp's mean and variance differ
    // from the warm start's by far more than anything else does)
    implied_stats(fp, mp, ev, ws, sm, theta, sig, lane);
    if (lane < mp.n_init) {
      const int k = mp.init_par[lane], e = mp.init_row[lane];
      double v = theta[k] + (s[e] - sig[e]);
      const double lb = fp.par_lbound[k];
      if (lb == lb && v < lb) v = lb;
      theta[k] = v;
    }
    __syncwarp();
  }
  double F = objective(theta, r);
  double dF_prev = 0.0;
  bool at_theta = true;   // the RAM scratch (Z, C, mu, sig) was last evaluated at theta: the Jacobian can reuse it
  double lam = 1e-4;
  int status = GWSEM_STATUS_ITERATION_LIMIT;
  int it = 0;
  const int max_iter = 200;
  for (it = 1; it <= max_iter && nbeta > 0; ++it) {
    marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Jm, b_row, nb, b_par, nbeta, false, lane, !at_theta);
    warp_forward_solve<32>(Lb, nb, Jm, nbeta, nbeta, lane);
    for (int k = lane; k < nbeta; k += 32) {
      double v = 0.0;
      for (int e = 0; e < nb; ++e) v += Jm[e * nbeta + k] * r[e];
      g[k] = v;
    }
    for (int idx = lane; idx < nbeta * nbeta; idx += 32) {
      const int a = idx / nbeta, b = idx % nbeta;
      double v = 0.0;
      for (int e = 0; e < nb; ++e) v += Jm[e * nbeta + a] * Jm[e * nbeta + b];
      H[idx] = v;
    }
    __syncwarp();
    bool improved = false;
    double Fc = F;
    for (int attempt = 0; attempt < 30; ++attempt) {
      for (int idx = lane; idx < nbeta * nbeta; idx += 32) {
        const int a = idx / nbeta, b = idx % nbeta;
        double v = H[idx];
        if (a == b) v += lam * fmax(H[idx], 1e-12);
        Hc[idx] = v;
      }
      __syncwarp();
      if (!warp_cholesky<32>(Hc, nbeta, lane)) { lam *= 10; continue; }
      if (lane == 0) {
        for (int i = 0; i < nbeta; ++i) {
          double v = g[i];
          for (int k = 0; k < i; ++k) v -= Hc[i * nbeta + k] * delta[k];
          delta[i] = v / Hc[i * nbeta + i];
        }
        for (int i = nbeta - 1; i >= 0; --i) {
          double v = delta[i];
          for (int k = i + 1; k < nbeta; ++k) v -= Hc[k * nbeta + i] * delta[k];
          delta[i] = v / Hc[i * nbeta + i];
        }
      }
      __syncwarp();
      for (int k = lane; k < P; k += 32) cand[k] = theta[k];
      __syncwarp();
      for (int c = lane; c < nbeta; c += 32) {
        const int k = b_par[c];
        double v = theta[k] + delta[c];
        const double lb = fp.par_lbound[k];
        if (lb == lb && v < lb) v = lb;
        cand[k] = v;
      }
      __syncwarp();
      Fc = objective(cand, rc);
      at_theta = false;
      if (isfinite(Fc) && Fc <= F) { improved = true; break; }
      lam *= 10;
    }
    if (!improved) {
      double gmax = 0.0;
      for (int k = 0; k < nbeta; ++k) gmax = fmax(gmax, fabs(g[k]));
      status = gmax < 1e-6 * (1 + F) ? GWSEM_STATUS_OK : GWSEM_STATUS_NONZERO_GRADIENT;
      break;
    }
    double step = 0.0;
    for (int c = 0; c < nbeta; ++c) {
      const int k = b_par[c];
      step = fmax(step, fabs(cand[k] - theta[k]) / (fabs(theta[k]) + 1e-3));
    }
    const double dF = F - Fc;
    __syncwarp();
    for (int k = lane; k < P; k += 32) theta[k] = cand[k];
    for (int e = lane; e < nb; e += 32) r[e] = rc[e];
    __syncwarp();
    F = Fc;
    at_theta = true;   // theta = cand, where the objective was just evaluated
    lam = fmax(lam / 10, 1e-12);
    // F is on the chi-square scale (the weight is the inverse covariance of the statistics), so a decrease dF moved the
    // estimate by about sqrt(2 dF) standard errors.  With the contraction kappa = sqrt(dF / previous dF) seen so far,
    // what is left after this step is sqrt(2 dF) kappa / (1 - kappa) standard errors: stop below 2e-6 (C3 cohort: kappa
    // around 0.01, three to four iterations; slowly contracting models run on to the oracle's own rule, which
    // marg_fit_kernel always uses).
    bool close_enough = false;
    if (it >= 2 && dF_prev > 0.0 && dF >= 0.0) {
      const double kappa = sqrt(dF / dF_prev);
      close_enough = kappa < 0.5 && sqrt(2.0 * dF) * kappa / (1.0 - kappa) <= 2e-6;
    }
    dF_prev = dF;
    if (step < 1e-10 || dF <= 1e-14 * (1 + F) || close_enough) { status = GWSEM_STATUS_OK; break; }
  }
  if (nbeta == 0) { status = GWSEM_STATUS_OK; it = 0; }
  if (it > max_iter) it = max_iter;

  lap(2);
  if (mp.diag && lane == 0) atomicAdd(&g_fast_clk[6], (unsigned long long)it);
  // ---- eliminated parameters: r_a = X (Lb^-1 r_b), then each one from its own statistic
  double *Da = mx + ms.Da, *Ja = mx + ms.Ja, *KV = mx + ms.KV;
  double* target = delta;   // na entries
  for (int i = lane; i < na; i += 32) {
    double ra = 0.0;
    for (int j = 0; j < nb; ++j) ra += X[i * nb + j] * r[j];
    target[i] = s[a_row[i]] - ra;
  }
  __syncwarp();
  for (int round = 0; round < 2; ++round) {
    marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Da, a_row, na, a_par, na, true, lane, round > 0 || !at_theta);
    for (int i = lane; i < na; i += 32) {
      const double d = Da[i];
      if (d != 0.0) theta[a_par[i]] += (target[i] - sig[a_row[i]]) / d;
    }
    __syncwarp();
  }
  lap(3);
  // Gamma positive definite <=> Gamma_bb and the Schur complement are: vcov(alpha) entries are read from S first
  // ---- parameter covariance
  marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Jm, b_row, nb, b_par, nbeta, false, lane);
  warp_forward_solve<32>(Lb, nb, Jm, nbeta, nbeta, lane);
  marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Ja, a_row, na, b_par, nbeta, false, lane, false);
  marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Da, a_row, na, a_par, na, true, lane, false);
  for (int idx = lane; idx < nbeta * nbeta; idx += 32) {
    const int a = idx / nbeta, b = idx % nbeta;
    double v = 0.0;
    for (int e = 0; e < nb; ++e) v += Jm[e * nbeta + a] * Jm[e * nbeta + b];
    H[idx] = v;
  }
  bool da_ok = true;
  for (int i = 0; i < na; ++i) da_ok = da_ok && Da[i] != 0.0;
  __syncwarp();
  const bool pd = da_ok && warp_cholesky<32>(H, nbeta, lane);
  if (pd) {
    chol_inverse(H, nbeta, Hc, lane);   // V = vcov(beta)
    // K = J_a,beta - X Jt (into Ja), KV = K V
    for (int e = lane; e < na * nbeta; e += 32) {
      const int i = e / nbeta, c = e % nbeta;
      double v = Ja[e];
      for (int k = 0; k < nb; ++k) v -= X[i * nb + k] * Jm[k * nbeta + c];
      Ja[e] = v;
    }
    __syncwarp();
    for (int e = lane; e < na * nbeta; e += 32) {
      const int i = e / nbeta, c = e % nbeta;
      double v = 0.0;
      for (int k = 0; k < nbeta; ++k) v += Ja[i * nbeta + k] * Hc[k * nbeta + c];
      KV[e] = v;
    }
    __syncwarp();
    for (int e = lane; e < nbeta * nbeta; e += 32) vcov[b_par[e / nbeta] * P + b_par[e % nbeta]] = Hc[e];
    for (int e = lane; e < na * nbeta; e += 32) {
      const int i = e / nbeta, c = e % nbeta;
      const double v = -KV[e] / Da[i];
      vcov[a_par[i] * P + b_par[c]] = v;
      vcov[b_par[c] * P + a_par[i]] = v;
    }
    for (int e = lane; e < na * na; e += 32) {
      const int x = e / na, y = e % na;
      double v = S[e];
      for (int c = 0; c < nbeta; ++c) v += KV[x * nbeta + c] * Ja[y * nbeta + c];
      vcov[a_par[x] * P + a_par[y]] = v / (Da[x] * Da[y]);
    }
  } else if (status == GWSEM_STATUS_OK) {
    status = GWSEM_STATUS_NOT_CONVEX;
  }
  __syncwarp();
  for (int e = lane; e < na; e += 32) Da[e] = fabs(S[e * na + e]);   // pivot scales of the Schur complement
  __syncwarp();
  if (!warp_cholesky<32>(S, na, lane, Da)) {
    // Gamma itself is not positive definite: no estimate is reported (as marg_fit_kernel)
    for (int k = lane; k < P * P; k += 32) vcov[k] = na_val;
    not_pd();
    return;
  }
  for (int k = lane; k < P; k += 32) est[k] = theta[k];
  if (lane == 0) {
    res.fit[snp] = F;
    res.status[snp] = status;
    res.iterations[snp] = it;
  }
  lap(4);
}

// The model description is a few KB of small tables that every step of the fit walks through (RAM cells, statistic
// kinds, elimination lists) plus three q0 x q0 constants; read from global memory their latency was the whole cost of
// the kernel (one SNP alone on the GPU: 80 k cycles per Gauss-Newton iteration).  A persistent CTA stages them in
// shared memory once and its warps then loop over SNPs.
struct FastTables {
  int total;   // bytes
};
template <typename T>
__device__ const T* stage_table(const T* src, int n, unsigned char* base, int& cursor) {
  T* dst = reinterpret_cast<T*>(base + cursor);
  if (src)
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  cursor += (n * (int)sizeof(T) + 7) & ~7;
  return src ? dst : nullptr;
}
__host__ __device__ inline int fast_table_bytes(const FitProgram& fp, const MargProgram& mp) {
  auto r8 = [](int b) { return (b + 7) & ~7; };
  const int nc = fp.n_cells, P = fp.P, ns = mp.n_stat, q0 = mp.q0;
  int b = 0;
  b += 2 * r8(4 * fp.q) + 5 * r8(4 * nc) + r8(8 * nc) + r8(4 * (P + 1)) + 3 * r8(8 * P) + r8(4 * ns) + r8(8 * ns) + r8(4 * mp.K);
  b += 4 * r8(4 * ns) + 2 * r8(4 * mp.n_a) + r8(4 * mp.n_b) + r8(4 * mp.n_beta) + r8(8 * mp.n1) + r8(8 * mp.npii) + 3 * r8(8 * q0 * q0);
  return b;
}

__global__ void __launch_bounds__(384, 1)
    marg_fit_fast_kernel(FitProgram fp_in, MargProgram mp_in, int n_snps, const double* __restrict__ summary,
                         const double* __restrict__ maf_in, gwsem_result res) {
  extern __shared__ double smem_d[];
  __shared__ __align__(16) unsigned char fp_raw[sizeof(FitProgram)];
  __shared__ __align__(16) unsigned char mp_raw[sizeof(MargProgram)];
  FitProgram& fp = *reinterpret_cast<FitProgram*>(fp_raw);
  MargProgram& mp = *reinterpret_cast<MargProgram*>(mp_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const WarpScratch ws = marg_ws(fp_in.p, fp_in.nv);
  const FastScratch ms = fast_layout(mp_in.nt, 2 + mp_in.J, mp_in.q0, mp_in.n_stat, fp_in.P, mp_in.n_a, mp_in.n_b, mp_in.n_beta);
  unsigned char* tab = reinterpret_cast<unsigned char*>(smem_d + (size_t)warps * (ws.total + ms.total));
  {
    int cur = 0;
    const int nc = fp_in.n_cells, P = fp_in.P, ns = mp_in.n_stat, q0 = mp_in.q0;
    const int32_t* stat_i = stage_table(fp_in.stat_i, fp_in.q, tab, cur);
    const int32_t* stat_j = stage_table(fp_in.stat_j, fp_in.q, tab, cur);
    const int32_t* cell_mat = stage_table(fp_in.cell_mat, nc, tab, cur);
    const int32_t* cell_row = stage_table(fp_in.cell_row, nc, tab, cur);
    const int32_t* cell_col = stage_table(fp_in.cell_col, nc, tab, cur);
    const int32_t* cell_par = stage_table(fp_in.cell_par, nc, tab, cur);
    const int32_t* par_cells;
    {
      int c2 = cur;   // every free cell is listed once: at most nc entries (the budget counts nc)
      par_cells = stage_table(fp_in.par_cells, min(nc, fp_in.par_cell_ptr[P]), tab, c2);
      cur += (4 * nc + 7) & ~7;
    }
    const double* cell_val = stage_table(fp_in.cell_val, nc, tab, cur);
    const int32_t* par_cell_ptr = stage_table(fp_in.par_cell_ptr, P + 1, tab, cur);
    const double* par_start = stage_table(fp_in.par_start, P, tab, cur);
    const double* par_lbound = stage_table(fp_in.par_lbound, P, tab, cur);
    const double* par_warm = stage_table(fp_in.par_warm, P, tab, cur);
    const int32_t* thr_par = stage_table(fp_in.thr_par, ns, tab, cur);
    const double* thr_val = stage_table(fp_in.thr_val, ns, tab, cur);
    const int32_t* exo_latent = stage_table(fp_in.exo_latent, mp_in.K, tab, cur);
    const int32_t* stat_kind = stage_table(mp_in.stat_kind, ns, tab, cur);
    const int32_t* stat_a = stage_table(mp_in.stat_a, ns, tab, cur);
    const int32_t* stat_b = stage_table(mp_in.stat_b, ns, tab, cur);
    const int32_t* stat_theta = stage_table(mp_in.stat_theta, ns, tab, cur);
    const int32_t* ea_par = stage_table(mp_in.elim_a_par, mp_in.n_a, tab, cur);
    const int32_t* ea_row = stage_table(mp_in.elim_a_row, mp_in.n_a, tab, cur);
    const int32_t* eb_row = stage_table(mp_in.elim_b_row, mp_in.n_b, tab, cur);
    const int32_t* eb_par = stage_table(mp_in.elim_b_par, mp_in.n_beta, tab, cur);
    const double* theta1 = stage_table(mp_in.theta1, mp_in.n1, tab, cur);
    const double* rho_ii = stage_table(mp_in.rho_ii, mp_in.npii, tab, cur);
    const double* B00 = stage_table(mp_in.B00, q0 * q0, tab, cur);
    const double* Wcc = stage_table(mp_in.Wcc, q0 * q0, tab, cur);
    const double* acov_cc = stage_table(mp_in.acov_cc, q0 * q0, tab, cur);
    if (threadIdx.x == 0) {
      fp = fp_in;
      mp = mp_in;
      fp.stat_i = stat_i; fp.stat_j = stat_j; fp.cell_mat = cell_mat; fp.cell_row = cell_row; fp.cell_col = cell_col;
      fp.cell_par = cell_par; fp.par_cells = par_cells; fp.cell_val = cell_val; fp.par_cell_ptr = par_cell_ptr;
      fp.par_start = par_start; fp.par_lbound = par_lbound; fp.par_warm = par_warm; fp.thr_par = thr_par; fp.thr_val = thr_val;
      fp.exo_latent = exo_latent;
      mp.stat_kind = stat_kind; mp.stat_a = stat_a; mp.stat_b = stat_b; mp.stat_theta = stat_theta;
      mp.elim_a_par = ea_par; mp.elim_a_row = ea_row; mp.elim_b_row = eb_row; mp.elim_b_par = eb_par;
      mp.theta1 = theta1; mp.rho_ii = rho_ii; mp.B00 = B00; mp.Wcc = Wcc; mp.acov_cc = acov_cc;
    }
  }
  __syncthreads();
  double* sm = smem_d + (size_t)warp * (ws.total + ms.total);
  for (int snp = blockIdx.x * warps + warp; snp < n_snps; snp += gridDim.x * warps) {
    fast_fit_one(fp, mp, ws, ms, snp, sm, summary, maf_in, res, lane);
    __syncwarp();
  }
}
#endif

__global__ void __launch_bounds__(GW_MF_LANES) marg_pattern_kernel(FitProgram fp, MargProgram mp, double* __restrict__ out) {
  extern __shared__ double smem_d[];
  const int lane = threadIdx.x;
  const int nt = mp.nt, ns = mp.n_stat, P = fp.P;
  const WarpScratch ws = marg_ws(fp.p, fp.nv);
  const MargScratch ms = marg_layout(nt, ns, P, 0);
  double *sm = smem_d, *mx = sm + ws.total;
  double *theta = mx + ms.theta, *sig = mx + ms.sig, *Jm = mx + ms.J;
  ModelEval<LANES> ev{fp, ws, sm, lane};
  for (int rep = 0; rep < 2; ++rep) {
    for (int k = lane; k < P; k += LANES)
      theta[k] = rep == 0 ? fp.par_start[k] + 0.1 + 0.013 * k : fp.par_start[k] * 1.3 + 0.2 - 0.011 * k;
    gsync<LANES>();
    marg_jacobian(fp, mp, ev, ws, sm, theta, sig, Jm, nullptr, ns, nullptr, P, false, lane);
    for (int e = lane; e < ns * P; e += LANES) out[rep * ns * P + e] = Jm[e];
    gsync<LANES>();
  }
}

}  // namespace GW_MF_NS
