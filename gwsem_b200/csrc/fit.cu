// Per-SNP WLS fit (cumulants summary statistics), one warp per SNP.
//
// What OpenMx does per loop iteration of the plan built by prepareComputePlan
// (reference R/model.R:187-194), restated for the GPU:
//   O1  naAction='omit' (people with a missing call leave every statistic), minVariance check
//   O2  sample covariance s (N-1) and ADF weight: Gamma[(ij),(kl)] = m_ijkl - s~_ij s~_kl (N)
//   O5  RAM expectation  Sigma = F (I-A)^-1 S (I-A)^-T F'
//   O6  F(theta) = (s - sigma)' W (s - sigma),  W = N Gamma^-1
//   O7  optimisation from the original start values (damped Gauss-Newton, analytic Jacobian,
//       lower bounds by projection)
//   O8  vcov = (J' W J)^-1
// The arithmetic mirrors oracle/fit_oracle.py step for step (same damping schedule), so the
// parity tests compare like with like.  All FP64.
#include "common.h"
#include "fit_device.cuh"
#include "fit_program.h"
#include "kernels.h"

namespace gwsem {

size_t fit_scratch_doubles(int p, int q, int nv, int P, int nf) { return scratch_layout(p, q, nv, P, nf).total; }

__device__ double central_moment(const FitProgram& fp, const int* vars, int k, const double* mu, const double* R,
                                 double inv_n) {
  const int p = fp.p, p1 = p + 1;
  double sum = 0.0;
  for (int mask = 0; mask < (1 << k); ++mask) {
    int sel[4] = {p, p, p, p};
    int c = 0;
    double coef = 1.0;
    for (int b = 0; b < k; ++b) {
      if ((mask >> b) & 1) sel[c++] = vars[b];
      else coef *= -mu[vars[b]];
    }
    double ev = 1.0;
    if (c) ev = R[fp.idx4[((sel[0] * p1 + sel[1]) * p1 + sel[2]) * p1 + sel[3]]] * inv_n;
    sum += coef * ev;
  }
  return sum;
}

// The model description is a few KB of small tables that every step of the fit walks through (the moment index, RAM
// cells, parameter lists).  Read from global memory their latency was most of the kernel's time (the lesson of
// marg_fit_fast_kernel): every CTA copies them into shared memory first.
template <typename T>
__device__ const T* stage_fit_table(const T* src, int n, unsigned char* base, int& cursor) {
  T* dst = reinterpret_cast<T*>(base + cursor);
  if (src)
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  cursor += (n * (int)sizeof(T) + 7) & ~7;
  return src ? dst : nullptr;
}
__host__ __device__ inline int fit_table_bytes(const FitProgram& fp) {
  auto r8 = [](int b) { return (b + 7) & ~7; };
  const int p1 = fp.p + 1;
  return r8(8 * fp.nf) + r8(4 * p1 * p1 * p1 * p1) + 2 * r8(4 * fp.q) + 5 * r8(4 * fp.n_cells) + r8(8 * fp.n_cells) +
         r8(4 * (fp.P + 1)) + 2 * r8(8 * fp.P);
}

template <int LANES>
__global__ void __launch_bounds__(128)
    fit_cumulants_kernel(FitProgram fp_in, int n_snps, const double* __restrict__ Rtab, const int32_t* __restrict__ n_rows,
                         const double* __restrict__ maf_in, gwsem_result res, int scratch_doubles) {
  extern __shared__ double smem_d[];
  __shared__ __align__(16) unsigned char fp_raw[sizeof(FitProgram)];
  // (one thread per SNP, LANES == 1: small models whose tables sit in L1 anyway; measured slower with the staging)
  const FitProgram& fp = LANES == 1 ? fp_in : *reinterpret_cast<const FitProgram*>(fp_raw);
  if (LANES > 1) {
    FitProgram& fps = *reinterpret_cast<FitProgram*>(fp_raw);
    unsigned char* tab = reinterpret_cast<unsigned char*>(smem_d + scratch_doubles);
    int cur = 0;
    const int p1 = fp_in.p + 1, nc = fp_in.n_cells, P = fp_in.P;
    const double* tot = stage_fit_table(fp_in.tot, fp_in.nf, tab, cur);
    const int32_t* idx4 = stage_fit_table(fp_in.idx4, p1 * p1 * p1 * p1, tab, cur);
    const int32_t* stat_i = stage_fit_table(fp_in.stat_i, fp_in.q, tab, cur);
    const int32_t* stat_j = stage_fit_table(fp_in.stat_j, fp_in.q, tab, cur);
    const int32_t* cell_mat = stage_fit_table(fp_in.cell_mat, nc, tab, cur);
    const int32_t* cell_row = stage_fit_table(fp_in.cell_row, nc, tab, cur);
    const int32_t* cell_col = stage_fit_table(fp_in.cell_col, nc, tab, cur);
    const int32_t* cell_par = stage_fit_table(fp_in.cell_par, nc, tab, cur);
    const int32_t* par_cells;
    {
      int c2 = cur;   // every free cell is listed once: at most nc entries
      par_cells = stage_fit_table(fp_in.par_cells, min(nc, fp_in.par_cell_ptr[P]), tab, c2);
      cur += (4 * nc + 7) & ~7;
    }
    const double* cell_val = stage_fit_table(fp_in.cell_val, nc, tab, cur);
    const int32_t* par_cell_ptr = stage_fit_table(fp_in.par_cell_ptr, P + 1, tab, cur);
    const double* par_start = stage_fit_table(fp_in.par_start, P, tab, cur);
    const double* par_lbound = stage_fit_table(fp_in.par_lbound, P, tab, cur);
    if (threadIdx.x == 0) {
      fps = fp_in;
      fps.tot = tot; fps.idx4 = idx4; fps.stat_i = stat_i; fps.stat_j = stat_j; fps.cell_mat = cell_mat; fps.cell_row = cell_row;
      fps.cell_col = cell_col; fps.cell_par = cell_par; fps.par_cells = par_cells; fps.cell_val = cell_val;
      fps.par_cell_ptr = par_cell_ptr; fps.par_start = par_start; fps.par_lbound = par_lbound;
    }
    __syncthreads();
  }
  const int group = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int snp = blockIdx.x * (blockDim.x / LANES) + group;
  if (snp >= n_snps) return;
  const int p = fp.p, q = fp.q, nv = fp.nv, P = fp.P, nf = fp.nf;
  const WarpScratch ws = scratch_layout(p, q, nv, P, nf);
  double* sm = smem_d + (size_t)group * (LANES == 1 ? (ws.total | 1) : ws.total);  // odd stride: no bank conflicts

  const int n = n_rows[snp];
  const double inv_n = 1.0 / (double)n;
  double* R = sm + ws.R;
  // raw moment table R[a][f] = sum over complete rows of g^a * h_f
  for (int e = lane; e < 5 * nf; e += LANES) R[e] = Rtab[(int64_t)snp * 5 * nf + e];
  gsync<LANES>();

  double* est = res.est + (int64_t)snp * P;
  double* vcov = res.vcov + (int64_t)snp * P * P;
  for (int k = lane; k < P; k += LANES) est[k] = fp.par_start[k];
  for (int k = lane; k < P * P; k += LANES) vcov[k] = __longlong_as_double((long long)GWSEM_NA_BITS);
  if (lane == 0) {
    res.maf[snp] = maf_in[snp];
    res.n_used[snp] = n;
    res.fit[snp] = __longlong_as_double((long long)GWSEM_NA_BITS);
    res.status[snp] = GWSEM_STATUS_NA;
    res.catch_code[snp] = GWSEM_CATCH_NONE;
    res.catch_var[snp] = -1;
    res.iterations[snp] = 0;
  }

  // ---- O1/O2: means, covariance, Gamma
  double* mu = sm + ws.mu;
  for (int v = lane; v < p; v += LANES) mu[v] = R[fp.idx4[((v * (p + 1) + p) * (p + 1) + p) * (p + 1) + p]] * inv_n;
  gsync<LANES>();
  double* c2 = sm + ws.c2;      // N-denominator central second moments
  double* sobs = sm + ws.sobs;  // N-1 denominator: the observed covariance statistics
  const double nm1 = (double)n / (double)(n - 1);
  for (int e = lane; e < q; e += LANES) {
    const int vars[2] = {fp.stat_j[e], fp.stat_i[e]};  // sorted ascending (j <= i)
    const double m = central_moment(fp, vars, 2, mu, R, inv_n);
    c2[e] = m;
    sobs[e] = m * nm1;
  }
  gsync<LANES>();
  int bad_var = -1;
  for (int v = 0; v < p; ++v) {
    const double var = sobs[vech_index(v, v, p)];
    if (!(var >= fp.min_variance)) { bad_var = v; break; }
  }
  if (n < 2) bad_var = 0;
  if (bad_var >= 0) {
    if (lane == 0) { res.catch_code[snp] = GWSEM_CATCH_MIN_VARIANCE; res.catch_var[snp] = bad_var; }
    return;
  }
  double* L = sm + ws.L;
  for (int idx = lane; idx < q * q; idx += LANES) {
    const int e1 = idx / q, e2 = idx % q;
    if (e2 > e1) { continue; }
    int vars[4] = {fp.stat_j[e1], fp.stat_i[e1], fp.stat_j[e2], fp.stat_i[e2]};
    // sort 4 ints
    for (int a = 1; a < 4; ++a) {
      const int key = vars[a];
      int b = a - 1;
      while (b >= 0 && vars[b] > key) { vars[b + 1] = vars[b]; --b; }
      vars[b + 1] = key;
    }
    const double m4 = central_moment(fp, vars, 4, mu, R, inv_n);
    const double gam = m4 - c2[e1] * c2[e2];
    L[e1 * q + e2] = gam;
    L[e2 * q + e1] = gam;
  }
  gsync<LANES>();
  // natural scale of Gamma[(ij),(ij)] is s_ii * s_jj; kept in `sig` until the optimiser needs it
  for (int e = lane; e < q; e += LANES)
    (sm + ws.sig)[e] = c2[vech_index(fp.stat_i[e], fp.stat_i[e], p)] * c2[vech_index(fp.stat_j[e], fp.stat_j[e], p)];
  gsync<LANES>();
  if (!warp_cholesky<LANES>(L, q, lane, sm + ws.sig)) {
    if (lane == 0) res.catch_code[snp] = GWSEM_CATCH_ACOV_NOT_PD;
    return;
  }

  // ---- O5-O7: damped Gauss-Newton on the whitened residual  r~ = L^-1 (s - sigma), F = n |r~|^2
  double *theta = sm + ws.theta, *cand = sm + ws.cand, *sig = sm + ws.sig, *r = sm + ws.r, *J = sm + ws.J;
  double *g = sm + ws.g, *H = sm + ws.H, *Hc = sm + ws.Hc, *delta = sm + ws.delta;
  ModelEval<LANES> ev{fp, ws, sm, lane};
  const double dn = (double)n;

  auto objective = [&](const double* th, double* resid) -> double {
    ev.implied(th, sig);
    for (int e = lane; e < q; e += LANES) resid[e] = sobs[e] - sig[e];
    gsync<LANES>();
    warp_forward_solve<LANES>(L, q, resid, 1, 1, lane);
    double f = 0.0;
    for (int e = 0; e < q; ++e) f += resid[e] * resid[e];
    return dn * f;
  };

  for (int k = lane; k < P; k += LANES) theta[k] = fp.par_start[k];
  gsync<LANES>();
  double F = objective(theta, r);
  double dF_prev = 0.0;
  bool at_theta = true;   // Z and C in the scratch were last evaluated at theta: the Jacobian can use them as they are
  double lam = 1e-4;
  int status = GWSEM_STATUS_ITERATION_LIMIT;
  int it = 0;
  const int max_iter = 200;
  for (it = 1; it <= max_iter; ++it) {
    if (!at_theta) ev.implied(theta, sig);
    ev.jacobian(J);
    warp_forward_solve<LANES>(L, q, J, P, P, lane);  // J~ = L^-1 J
    for (int k = lane; k < P; k += LANES) {
      double v = 0.0;
      for (int e = 0; e < q; ++e) v += J[e * P + k] * r[e];
      g[k] = dn * v;
    }
    for (int idx = lane; idx < P * P; idx += LANES) {
      const int a = idx / P, b = idx % P;
      double v = 0.0;
      for (int e = 0; e < q; ++e) v += J[e * P + a] * J[e * P + b];
      H[idx] = dn * v;
    }
    gsync<LANES>();
    bool improved = false;
    double Fc = F;
    for (int attempt = 0; attempt < 30; ++attempt) {
      for (int idx = lane; idx < P * P; idx += LANES) {
        const int a = idx / P, b = idx % P;
        double v = H[idx];
        if (a == b) v += lam * fmax(H[idx], 1e-12);
        Hc[idx] = v;
      }
      gsync<LANES>();
      if (!warp_cholesky<LANES>(Hc, P, lane)) { lam *= 10; continue; }
      if (lane == 0) {  // solve Hc Hc' delta = g
        for (int i = 0; i < P; ++i) {
          double v = g[i];
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
      Fc = objective(cand, sm + ws.rc);
      at_theta = false;
      if (isfinite(Fc) && Fc <= F) { improved = true; break; }
      lam *= 10;
    }
    if (!improved) {
      double gmax = 0.0;
      for (int k = 0; k < P; ++k) gmax = fmax(gmax, fabs(g[k]));
      status = gmax < 1e-6 * (1 + F) ? GWSEM_STATUS_OK : GWSEM_STATUS_NONZERO_GRADIENT;
      break;
    }
    double step = 0.0;
    for (int k = 0; k < P; ++k) step = fmax(step, fabs(cand[k] - theta[k]) / (fabs(theta[k]) + 1e-3));
    const double dF = F - Fc;
    gsync<LANES>();
    for (int k = lane; k < P; k += LANES) theta[k] = cand[k];
    for (int e = lane; e < q; e += LANES) r[e] = (sm + ws.rc)[e];
    gsync<LANES>();
    F = Fc;
    at_theta = true;   // theta = cand, where the objective was just evaluated
    lam = fmax(lam / 10, 1e-12);
    // F is on the chi-square scale: a decrease dF moved the estimate by about sqrt(2 dF) standard errors; with the
    // contraction kappa = sqrt(dF / previous dF) seen so far, sqrt(2 dF) kappa / (1 - kappa) standard errors are left.
    // Stop below 2e-6 (as marg_fit_fast_kernel does); slowly contracting fits run on to the oracle's own rule.
    bool close_enough = false;
    if (it >= 2 && dF_prev > 0.0 && dF >= 0.0) {
      const double kappa = sqrt(dF / dF_prev);
      close_enough = kappa < 0.5 && sqrt(2.0 * dF) * kappa / (1.0 - kappa) <= 2e-6;
    }
    dF_prev = dF;
    if (step < 1e-10 || dF <= 1e-14 * (1 + F) || close_enough) { status = GWSEM_STATUS_OK; break; }
  }
  if (it > max_iter) it = max_iter;

  // ---- O8: vcov = (J' W J)^-1 at the solution
  if (!at_theta) ev.implied(theta, sig);
  ev.jacobian(J);
  warp_forward_solve<LANES>(L, q, J, P, P, lane);
  for (int idx = lane; idx < P * P; idx += LANES) {
    const int a = idx / P, b = idx % P;
    double v = 0.0;
    for (int e = 0; e < q; ++e) v += J[e * P + a] * J[e * P + b];
    H[idx] = dn * v;
  }
  gsync<LANES>();
  const bool pd = warp_cholesky<LANES>(H, P, lane);
  if (pd) {
    // inverse via solves against the identity: column c handled by lane c
    for (int c = lane; c < P; c += LANES) {
      double* x = Hc + c;  // column c of Hc (stride P)
      for (int i = 0; i < P; ++i) {
        double v = (i == c) ? 1.0 : 0.0;
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
    status = GWSEM_STATUS_NOT_CONVEX;   // the information matrix is not positive definite at the solution
  }
  for (int k = lane; k < P; k += LANES) est[k] = theta[k];
  if (lane == 0) {
    res.fit[snp] = F;
    res.status[snp] = status;
    res.iterations[snp] = it;
  }
}

void launch_fit_cumulants(gwsem_engine* e, const FitProgram& fp, int n_snps, const double* d_R, const int32_t* d_n,
                          const double* d_maf, const gwsem_result& d_res) {
  if (n_snps <= 0) return;
  const size_t per_snp = fit_scratch_doubles(fp.p, fp.q, fp.nv, fp.P, fp.nf) * sizeof(double);
  const size_t tables = (size_t)fit_table_bytes(fp) + 16;
  e->prof_begin(kFamFit);
  if (per_snp * 64 <= 96 * 1024) {
    // small model: one thread per SNP, each with its own shared-memory slab
    const int threads = 64;
    const size_t scratch = (per_snp + 8) * threads;
    const size_t smem = scratch;   // (no staged tables in this variant: they would only cost occupancy)
    GW_CUDA(cudaFuncSetAttribute(fit_cumulants_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fit_cumulants_kernel<1><<<(n_snps + threads - 1) / threads, threads, smem, e->stream>>>(fp, n_snps, d_R, d_n, d_maf, d_res,
                                                                                            (int)(scratch / sizeof(double)));
  } else {
    int warps = 4;
    while (warps > 1 && per_snp * warps + tables > 200 * 1024) warps >>= 1;
    if (per_snp * warps + tables > 227 * 1024) fail("model too large for the per-SNP fit kernel (%zu bytes of scratch)", per_snp);
    const size_t scratch = per_snp * warps;
    const size_t smem = scratch + tables;
    GW_CUDA(cudaFuncSetAttribute(fit_cumulants_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fit_cumulants_kernel<32><<<(n_snps + warps - 1) / warps, warps * 32, smem, e->stream>>>(fp, n_snps, d_R, d_n, d_maf, d_res,
                                                                                             (int)(scratch / sizeof(double)));
  }
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamFit);
  e->launches++;
}

}  // namespace gwsem
