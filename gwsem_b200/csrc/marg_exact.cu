// WLS "marginals" summary statistics recomputed in full for one SNP:
// the generic path behind marg_stats_kernel's fast one.  OpenMx re-estimates every marginal statistic on
// the rows that survive naAction = 'omit' for the SNP at hand (reference R/model.R:435-436), so whenever
// the retained people differ from the model's people (a missing call), or a variable / covariate itself
// carries the genotype (snp or a gxe product `snp_<moderator>` as manifest variable, R/model.R:427-434;
// snp as definition variable of a multi-item buildItem model, R/model.R:560-576), nothing is
// SNP-independent and this kernel does the whole of stage 1, stage 2 and Muthen's score products:
//   A  OLS of every continuous variable on its covariates (normal equations from one Gram pass, residual
//      sum of squares from a second), raw variances for the minVariance check
//   B  ordered probit of every ordinal variable: Newton-Raphson with the exact Hessian
//   C  one correlation per pair by Newton on the score with its exact derivative: Pearson, polyserial,
//      polychoric (bivariate normal rectangles by Gauss-Legendre in the Drezner-Wesolowsky form)
//   D  per-person score rows at the estimates -> B = sum sc sc', A21 = sum sc2 * d loglik2 / d theta1
// One CTA per SNP; a tile of T people at a time is staged in shared memory (thread = person), then
// every thread owns output entries and sums the tile's people in a fixed order (deterministic,
// independent of the batch).  The record (n, catch, theta, A21, B) goes to marg_fit_kernel.
// Mirrors oracle/marginals_oracle.py (marginals_stats) term for term.
#include <algorithm>

#include "common.h"
#include "kernels.h"
#include "marginals.h"

namespace gwsem {

namespace {

constexpr double kZmax = 12.0;   // stands in for +-infinity thresholds (as the oracle)
constexpr double kInvSqrt2Pi = 0.3989422804014327;
constexpr double kInv2Pi = 0.15915494309189535;

__device__ __forceinline__ double nphi(double z) { return kInvSqrt2Pi * exp(-0.5 * z * z); }
// Phi(a1) - Phi(a0), taken in the tail where both are small
__device__ __forceinline__ double ndiff(double a1, double a0) {
  return a0 > 0.0 ? normcdf(-a0) - normcdf(-a1) : normcdf(a1) - normcdf(a0);
}

// shared-memory layout (doubles)
struct ExLayout {
  int th, rho_ok, acc, part, g, x, y, st, sc, ps, etab, gtab, total;
};
__host__ __device__ inline int ex_acc_entries(const ExactProgram& ex) {
  int w = 1 + ex.K, hb = 0;
  for (int j = 0; j < ex.p; ++j) {
    if (ex.ncat[j] == 0) ++w;
    else { const int d = ex.voff[j + 1] - ex.voff[j]; hb += d + d * d; }
  }
  int a21 = 0;
  for (int b = 0; b < ex.p; ++b)
    for (int a = b + 1; a < ex.p; ++a) a21 += (ex.voff[a + 1] - ex.voff[a]) + (ex.voff[b + 1] - ex.voff[b]);
  int m = w * (w + 1) / 2;
  if (hb > m) m = hb;
  const int d = ex.nt * (ex.nt + 1) / 2 + a21;
  if (d > m) m = d;
  if (2 * ex.npairs > m) m = 2 * ex.npairs;
  return m;
}
__host__ __device__ inline ExLayout ex_layout(const ExactProgram& ex) {
  ExLayout l;
  int o = 0;
  auto take = [&](int n) { int at = o; o += n; return at; };
  const int T = ex.T;
  l.th = take(ex.nt);
  l.rho_ok = take(ex.npairs > 0 ? ex.npairs : 1);
  l.acc = take(ex_acc_entries(ex));
  l.part = take(1024);   // partial sums of the chunked accumulations (entries x chunks <= 2 x threads)
  l.g = take(T);
  l.x = take(T * ex.Kp);
  l.y = take(T * ex.p);
  l.st = take(T * ex.p * 4);
  const int w = 1 + ex.K + ex.p;
  l.sc = take(T * (ex.nt > w ? ex.nt : w));
  l.ps = take(T * (ex.npairs > 0 ? ex.npairs : 1) * 4);
  l.etab = take((ex_acc_entries(ex) + 1) / 2 + 1);   // int32 entry tables (Hessian entries; then B entries and A21 entries)
  l.gtab = take((ex.npairs > 0 ? ex.npairs : 1) * (1 + 3 * 32));
  l.total = o;
  return l;
}

// The small lookup tables of the program, copied to shared memory once per CTA: indexed dynamically in every inner
// loop, they would otherwise live in per-thread local memory (kernel parameters and local arrays alike), which
// the shared-memory carve-out leaves almost no L1 for.
struct ExShared {
  int voff[kExMaxVars + 1], ncat[kExMaxVars], nx[kExMaxVars], xidx[kExMaxVars][kExMaxCov];
  int ord_of[kExMaxVars], hoff[kExMaxVars + 1], cont_of[kExMaxVars], n_ord, n_cont;
  unsigned char powner[kExMaxVars * (kExMaxVars - 1) / 2];
};

struct Pair { int a, b; };   // a > b
__device__ __forceinline__ Pair pair_of(int k, int p) {
  int b = 0;
  while (k >= p - 1 - b) { k -= p - 1 - b; ++b; }
  return Pair{b + 1 + k, b};
}

// Standardised quantities of variable j for one person at the current theta:
//   continuous: s.a = u = e / sd, s.b = e ;  ordinal: s.a = z1, s.b = z0 (sentinels +-12 at the open ends)
struct VState { double a, b; };
__device__ __forceinline__ VState var_state(const ExShared& ex, const double* th, int j, const double* xr, double yv) {
  const int o = ex.voff[j], nx = ex.nx[j];
  VState s;
  if (ex.ncat[j] == 0) {
    double e = yv - th[o];
    for (int m = 0; m < nx; ++m) e -= th[o + 1 + m] * xr[ex.xidx[j][m]];
    s.b = e;
    s.a = e * rsqrt(th[o + 1 + nx]);
  } else {
    const int C = ex.ncat[j], c = (int)yv;
    double eta = 0.0;
    for (int m = 0; m < nx; ++m) eta += th[o + C - 1 + m] * xr[ex.xidx[j][m]];
    s.a = (c < C - 1 ? th[o + c] : kZmax) - eta;
    s.b = (c > 0 ? th[o + c - 1] : -kZmax) - eta;
  }
  return s;
}

// Rectangle probability of the standard bivariate normal and what the polychoric score needs:
//   P = Phi2-rectangle([a0,a1] x [b0,b1]; rho), Dn = sum of +-phi2 at the finite corners, dDn = d Dn / d rho
struct Poly { double P, Dn, dDn; };
constexpr int kGlMax = 32, kGlStride = 1 + 3 * kGlMax;
// tb: {nodes, then per node sin t, 1 / (2 cos^2 t), weight * asin(rho) / 2} (marg_exact_kernel's fill_gl_tables)
__device__ Poly polychoric_terms(double a1, double a0, double b1, double b0, bool au, bool al, bool bu, bool bl,
                                 double rho, const double* tb) {
  Poly r;
  const double s2 = 1.0 - rho * rho, is2 = 1.0 / s2, nrm = kInv2Pi * rsqrt(s2);
  const int gln = (int)tb[0];
  double integral = 0.0;
  r.Dn = 0.0;
  r.dDn = 0.0;
  const double hs[2] = {a1, a0}, ks[2] = {b1, b0};
  const bool hf[2] = {au, al}, kf[2] = {bu, bl};
#pragma unroll
  for (int ci = 0; ci < 2; ++ci)
#pragma unroll
    for (int cj = 0; cj < 2; ++cj) {
      if (!hf[ci] || !kf[cj]) continue;   // a corner on an open bound: no density, no correlation-dependent mass
      const double h = hs[ci], k = ks[cj], sgn = (ci == cj) ? 1.0 : -1.0;
      const double hk2 = 2.0 * h * k, hh = h * h + k * k;
      double acc = 0.0;
      for (int i = 0; i < gln; ++i) acc += tb[3 + 3 * i] * exp((hk2 * tb[1 + 3 * i] - hh) * tb[2 + 3 * i]);
      integral += sgn * acc;
      const double q = hh - rho * hk2;
      const double dens = nrm * exp(-0.5 * q * is2);
      r.Dn += sgn * dens;
      r.dDn += sgn * dens * (rho * is2 + (0.5 * hk2 * s2 - rho * q) * is2 * is2);
    }
  r.P = ndiff(a1, a0) * ndiff(b1, b0) + integral * kInv2Pi;
  if (!(r.P > 1e-300)) r.P = 1e-300;
  return r;
}

struct PairEval {
  double psi, dpsi;   // d loglik_i / d rho and its rho-derivative
  double a0, a1;      // side a: ordinal (d/dz1, d/dz0); continuous (d/du, unused)
  double b0, b1;      // side b
};

// One person's contribution to pair (a > b); sa / sb from var_state.  need_sides: also the derivatives with
// respect to the two variables' standardised quantities (phase D only).
__device__ PairEval pair_eval(const ExShared& ex, int a, int b, VState sa, VState sb, int ca, int cb, double rho,
                              const double* gl /* this pair's node table */, bool need_sides) {
  PairEval r;
  r.a0 = r.a1 = r.b0 = r.b1 = 0.0;
  const bool oa = ex.ncat[a] > 0, ob = ex.ncat[b] > 0;
  const double s2 = 1.0 - rho * rho;
  if (!oa && !ob) {
    const double u = sa.a, v = sb.a, is2 = 1.0 / s2;
    const double num = rho * s2 + u * v * (1.0 + rho * rho) - rho * (u * u + v * v);
    r.psi = num * is2 * is2;
    const double dnum = 1.0 - 3.0 * rho * rho + 2.0 * rho * u * v - (u * u + v * v);
    r.dpsi = dnum * is2 * is2 + 4.0 * rho * num * is2 * is2 * is2;
    r.a0 = -(u - rho * v) * is2;
    r.b0 = -(v - rho * u) * is2;
    return r;
  }
  if (oa != ob) {   // polyserial: c continuous, o ordinal
    const int io = oa ? a : b;
    const VState so = oa ? sa : sb, scn = oa ? sb : sa;
    const int co = oa ? ca : cb, C = ex.ncat[io];
    const double u = scn.a, z1 = so.a, z0 = so.b;
    const double is = rsqrt(s2), is3 = is * is * is;
    const double x1 = (z1 - rho * u) * is, x0 = (z0 - rho * u) * is;
    const double f1 = co < C - 1 ? nphi(x1) : 0.0, f0 = co > 0 ? nphi(x0) : 0.0;
    double pi = ndiff(x1, x0);
    if (!(pi > 1e-300)) pi = 1e-300;
    const double ip = 1.0 / pi;
    const double g1 = (rho * z1 - u) * is3, g0 = (rho * z0 - u) * is3;   // d x / d rho
    const double r3 = 3.0 * rho * is * is;
    const double h1 = r3 * g1 + z1 * is3, h0 = r3 * g0 + z0 * is3;       // d^2 x / d rho^2
    r.psi = (f1 * g1 - f0 * g0) * ip;
    r.dpsi = (f1 * (h1 - x1 * g1 * g1) - f0 * (h0 - x0 * g0 * g0)) * ip - r.psi * r.psi;
    const double dz1 = f1 * is * ip, dz0 = -f0 * is * ip, du = -u - rho * is * (f1 - f0) * ip;
    if (oa) { r.a0 = dz1; r.a1 = dz0; r.b0 = du; }
    else { r.b0 = dz1; r.b1 = dz0; r.a0 = du; }
    return r;
  }
  // polychoric
  const int Ca = ex.ncat[a], Cb = ex.ncat[b];
  const bool au = ca < Ca - 1, al = ca > 0, bu = cb < Cb - 1, bl = cb > 0;
  const Poly t = polychoric_terms(sa.a, sa.b, sb.a, sb.b, au, al, bu, bl, rho, gl);
  const double ip = 1.0 / t.P;
  r.psi = t.Dn * ip;
  r.dpsi = t.dDn * ip - r.psi * r.psi;
  if (need_sides) {
    const double is = rsqrt(s2);
    auto edge = [&](double za, double zb1, double zb0) { return nphi(za) * ndiff((zb1 - rho * za) * is, (zb0 - rho * za) * is); };
    r.a0 = au ? edge(sa.a, sb.a, sb.b) * ip : 0.0;
    r.a1 = al ? -edge(sa.b, sb.a, sb.b) * ip : 0.0;
    r.b0 = bu ? edge(sb.a, sa.a, sa.b) * ip : 0.0;
    r.b1 = bl ? -edge(sb.b, sa.a, sa.b) * ip : 0.0;
  }
  return r;
}

// d loglik2_i / d theta_c for stage-1 parameter `l` (local index in variable v's block), from the two staged
// scalars of that side (already multiplied by the pair's score): ordinal (s0 = psi d/dz1, s1 = psi d/dz0),
// continuous (s0 = psi * (-d/du / sd), s1 = psi * (-1/(2 var) - d/du * u / (2 var))).
__device__ __forceinline__ double side_entry(const ExShared& ex, int v, int l, double s0, double s1, double yv,
                                             const double* xr) {
  const int nx = ex.nx[v];
  if (ex.ncat[v] > 0) {
    const int C1 = ex.ncat[v] - 1, c = (int)yv;
    if (l < C1) return (c == l ? s0 : 0.0) + (c - 1 == l ? s1 : 0.0);
    return -xr[ex.xidx[v][l - C1]] * (s0 + s1);
  }
  if (l == 0) return s0;
  if (l <= nx) return xr[ex.xidx[v][l - 1]] * s0;
  return s1;
}

// In-place Cholesky solve by one thread (d <= ~25): H (d x d, row-major, symmetric) x = g -> g.  false if not PD.
__device__ bool thread_chol_solve(double* H, double* g, int d) {
  for (int j = 0; j < d; ++j) {
    double v = H[j * d + j];
    for (int k = 0; k < j; ++k) v -= H[j * d + k] * H[j * d + k];
    if (!(v > 1e-12 * fabs(H[j * d + j])) || !isfinite(v)) return false;
    v = sqrt(v);
    H[j * d + j] = v;
    for (int i = j + 1; i < d; ++i) {
      double w = H[i * d + j];
      for (int k = 0; k < j; ++k) w -= H[i * d + k] * H[j * d + k];
      H[i * d + j] = w / v;
    }
  }
  for (int i = 0; i < d; ++i) {
    double v = g[i];
    for (int k = 0; k < i; ++k) v -= H[i * d + k] * g[k];
    g[i] = v / H[i * d + i];
  }
  for (int i = d - 1; i >= 0; --i) {
    double v = g[i];
    for (int k = i + 1; k < d; ++k) v -= H[k * d + i] * g[k];
    g[i] = v / H[i * d + i];
  }
  return true;
}

__global__ void marg_exact_kernel(ExactProgram ex, const uint8_t* __restrict__ rows, int64_t pitch, int n_src,
                                  const uint8_t* __restrict__ inc8, const int32_t* __restrict__ counts, int plink1,
                                  const double* __restrict__ geno, int n_pad, double* __restrict__ records) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int flag_sh, nlive_sh;
  __shared__ ExShared S;
  const int snp = blockIdx.x, tid = threadIdx.x, NT = blockDim.x, T = ex.T, pt = tid % T, sub = tid / T, R = NT / T, p = ex.p, K = ex.K, Kp = ex.Kp, nt = ex.nt, np = ex.npairs;
  double* rec = records + (int64_t)snp * ex.full_stride;
  // only_missing: SNPs without a missing call keep the fast path's record
  if (ex.only_missing && counts && counts[8 * (int64_t)snp + 2] == 0) {
    if (tid == 0) rec[0] = -1.0;
    return;
  }
  if (tid == 0) {
    S.n_ord = S.n_cont = 0;
    S.hoff[0] = 0;
    for (int j = 0; j <= ex.p; ++j) S.voff[j] = ex.voff[j];
    for (int j = 0; j < ex.p; ++j) {
      S.ncat[j] = ex.ncat[j];
      S.nx[j] = ex.nx[j];
      for (int m = 0; m < ex.nx[j]; ++m) S.xidx[j][m] = ex.xidx[j][m];
      if (ex.ncat[j] == 0) S.cont_of[S.n_cont++] = j;
      else {
        const int d = ex.voff[j + 1] - ex.voff[j];
        S.ord_of[S.n_ord] = j;
        S.hoff[S.n_ord + 1] = S.hoff[S.n_ord] + d + d * d;
        ++S.n_ord;
      }
    }
    int n_poly = 0, n_other = 0;
    const int R0 = blockDim.x / ex.T;
    for (int k2 = 0, b2 = 0; b2 < ex.p; ++b2)
      for (int a2 = b2 + 1; a2 < ex.p; ++a2, ++k2) {
        const bool poly = ex.ncat[a2] > 0 && ex.ncat[b2] > 0;
        S.powner[k2] = (unsigned char)((poly ? n_poly++ : n_other++) % R0);
      }
  }
  __syncthreads();
  const ExLayout L = ex_layout(ex);
  double *th = sm + L.th, *rho_ok = sm + L.rho_ok, *acc = sm + L.acc, *gs = sm + L.g, *xs = sm + L.x, *ys = sm + L.y;
  double *st = sm + L.st, *sc = sm + L.sc, *ps = sm + L.ps;
  const int n_acc = ex_acc_entries(ex);
  const uint8_t* row = rows ? rows + (int64_t)snp * pitch : nullptr;
  const double* grow = geno ? geno + (int64_t)snp * n_pad : nullptr;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);

  // stage one tile of people: genotype, covariates, variables (NaN genotype = not retained)
  auto load_tile = [&](int t0) {
    if (sub == 0) {
    const int i = t0 + pt;
    double g = nan;
    if (i < n_src && inc8[i]) {
      if (grow) g = grow[i];
      else {
        unsigned code = (row[i >> 2] >> (2 * (i & 3))) & 3u;
        if (plink1) code = (0x1Eu >> (2 * code)) & 3u;
        if (code != 3u) g = (double)code;
      }
    }
    gs[pt] = g;
    const bool live = g == g;
    for (int k = 0; k < K; ++k) {
      double x = 0.0;
      if (live) {
        const int kd = ex.xkind[k];
        x = kd == 1 ? g : (kd == 2 ? g * ex.X[(int64_t)i * Kp + k] : ex.X[(int64_t)i * Kp + k]);
      }
      xs[pt * Kp + k] = x;
    }
    for (int j = 0; j < p; ++j) {
      double y = 0.0;
      if (live) {
        const int kd = ex.vkind[j];
        y = kd == 2 ? g : (kd == 3 ? g * ex.V[(int64_t)i * p + j] : ex.V[(int64_t)i * p + j]);
      }
      ys[pt * p + j] = y;
    }
    }
    __syncthreads();
  };
  auto zero_acc = [&]() {
    for (int e = tid; e < n_acc; e += NT) acc[e] = 0.0;
  };
  // acc[dst(e)] += sum over the tile's people of f(e, t) for e < n_e: when there are fewer entries than threads the
  // people are split into chunks (one thread per entry and chunk), partial sums combined in chunk order
  double* part = sm + L.part;
  auto chunked = [&](int n_e, auto f, auto dst) {
    int CH = n_e > 0 ? NT / n_e : 1;
    if (CH < 1) CH = 1;
    if (CH > 8) CH = 8;
    if (n_e * CH > 1024) CH = 1;
    if (CH == 1) {
      for (int e = tid; e < n_e; e += NT) {
        double v = 0.0;
        for (int t = 0; t < T; ++t) v += f(e, t);
        dst(e, v);
      }
      return;
    }
    for (int w = tid; w < n_e * CH; w += NT) {
      const int e = w % n_e, c = w / n_e;
      double v = 0.0;
      for (int t = c * T / CH; t < (c + 1) * T / CH; ++t) v += f(e, t);
      part[w] = v;
    }
    __syncthreads();
    for (int e = tid; e < n_e; e += NT) {
      double v = 0.0;
      for (int c = 0; c < CH; ++c) v += part[c * n_e + e];
      dst(e, v);
    }
  };
  auto finish = [&](int code, int var) {
    if (tid == 0) { rec[1] = (double)code; rec[2] = (double)var; }
  };

  long long clk0 = clock64();
  double diag[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // cycles of phases A..D, Newton iterations of B and C (development aid)
  auto lap = [&](int slot) { const long long c = clock64(); diag[slot] += (double)(c - clk0); clk0 = c; };
  for (int e = tid; e < nt; e += NT) th[e] = ex.theta0[e];
  for (int e = tid; e < ex.full_stride; e += NT) rec[e] = 0.0;
  if (tid == 0) { flag_sh = 0; nlive_sh = 0; }
  zero_acc();
  __syncthreads();

  // ---------------- phase A: Gram of w = [1, x_1..x_K, continuous variables] over the retained people
  const int n_cont = S.n_cont;
  const int* cont_of = S.cont_of;
  const int W = 1 + K + n_cont, nW = W * (W + 1) / 2;
  int n_live_local = 0;
  for (int t0 = 0; t0 < n_src; t0 += T) {
    load_tile(t0);
    if (sub == 0) {
      const bool live = gs[pt] == gs[pt];
      n_live_local += live;
      double* w = sc + pt * W;
      w[0] = live ? 1.0 : 0.0;
      for (int k = 0; k < K; ++k) w[1 + k] = xs[pt * Kp + k];
      for (int c = 0; c < n_cont; ++c) w[1 + K + c] = ys[pt * p + cont_of[c]];
    }
    __syncthreads();
    for (int e = tid; e < nW; e += NT) {
      int a = 0, rem = e;
      while (rem >= W - a) { rem -= W - a; ++a; }
      const int b = a + rem;
      double v = 0.0;
      for (int t = 0; t < T; ++t) v = fma(sc[t * W + a], sc[t * W + b], v);
      acc[e] += v;
    }
    __syncthreads();
  }
  atomicAdd(&nlive_sh, n_live_local);
  __syncthreads();
  const int n = nlive_sh;
  const double dn = (double)n;
  if (tid == 0) rec[0] = dn;
  auto gram = [&](int a, int b) -> double {   // a <= b
    if (a > b) { const int t2 = a; a = b; b = t2; }
    return acc[a * W - a * (a - 1) / 2 + (b - a)];
  };
  // minVariance (reference R/model.R:435-436 mxData(minVariance=)): the N-1 variance of the genotype when it is a
  // predictor, then of every continuous manifest variable in order
  if (n < 2) { finish(GWSEM_CATCH_MIN_VARIANCE, ex.snp_exo ? -2 : 0); return; }
  if (tid == 0) {
    int bad = -1;
    if (ex.snp_exo) {
      for (int k = 0; k < K && bad == -1; ++k)
        if (ex.xkind[k] == 1) {
          const double v = (gram(1 + k, 1 + k) - gram(0, 1 + k) * gram(0, 1 + k) / dn) / (dn - 1.0);
          if (!(v >= ex.min_variance)) bad = -2;
        }
    }
    for (int c = 0; c < n_cont && bad == -1; ++c) {
      const double v = (gram(1 + K + c, 1 + K + c) - gram(0, 1 + K + c) * gram(0, 1 + K + c) / dn) / (dn - 1.0);
      if (!(v >= ex.min_variance)) bad = cont_of[c];
    }
    if (bad != -1) { flag_sh = 1; rec[1] = (double)GWSEM_CATCH_MIN_VARIANCE; rec[2] = (double)bad; }
  }
  __syncthreads();
  if (flag_sh) return;
  // OLS of each continuous variable (thread c solves variable cont_of[c]); scratch after the Gram in acc
  if (tid < n_cont) {
    const int j = cont_of[tid], d = 1 + S.nx[j];
    double* H = sc + tid * ((kExMaxCov + 1) * (kExMaxCov + 2));   // scratch: the tile buffers are idle here
    double* gv = H + (kExMaxCov + 1) * (kExMaxCov + 1);
    auto col = [&](int l) { return l == 0 ? 0 : 1 + S.xidx[j][l - 1]; };
    for (int a = 0; a < d; ++a) {
      gv[a] = gram(col(a), 1 + K + tid);
      for (int b = 0; b < d; ++b) H[a * d + b] = gram(col(a), col(b));
    }
    if (!thread_chol_solve(H, gv, d)) atomicExch(&flag_sh, 2);
    for (int a = 0; a < d; ++a) th[S.voff[j] + a] = gv[a];
    th[S.voff[j] + d] = 1.0;   // variance: second pass
  }
  __syncthreads();
  if (flag_sh) { finish(GWSEM_CATCH_ACOV_NOT_PD, -1); return; }
  // residual sums of squares
  zero_acc();
  __syncthreads();
  for (int t0 = 0; t0 < n_src; t0 += T) {
    load_tile(t0);
    const bool live = gs[pt] == gs[pt];
    for (int c = sub; c < n_cont; c += R) {
      double e2 = 0.0;
      if (live) {
        const VState s = var_state(S, th, cont_of[c], xs + pt * Kp, ys[pt * p + cont_of[c]]);
        e2 = s.b * s.b;
      }
      sc[pt * n_cont + c] = e2;
    }
    __syncthreads();
    for (int e = tid; e < n_cont; e += NT) {
      double v = 0.0;
      for (int t = 0; t < T; ++t) v += sc[t * n_cont + e];
      acc[e] += v;
    }
    __syncthreads();
  }
  if (tid < n_cont) {
    const int j = cont_of[tid];
    const double var = acc[tid] / dn;
    th[S.voff[j] + 1 + S.nx[j]] = var;
    if (!(var > 0.0)) atomicExch(&flag_sh, 2);
  }
  __syncthreads();
  if (flag_sh) { finish(GWSEM_CATCH_ACOV_NOT_PD, -1); return; }

  lap(0);
  // ---------------- phase B: ordered probit of every ordinal variable, Newton with the exact Hessian
  const int n_ord = S.n_ord;
  const int *ord_of = S.ord_of, *hoff = S.hoff;
  // entries of every ordinal variable's gradient and Hessian (lower triangle): (variable, a, b), b = 255 for the gradient
  int* htab = reinterpret_cast<int*>(sm + L.etab);
  int n_h = 0;
  for (int o = 0; o < n_ord; ++o) {
    const int d = S.voff[ord_of[o] + 1] - S.voff[ord_of[o]];
    n_h += d + d * (d + 1) / 2;
  }
  for (int e = tid; e < n_h; e += NT) {
    int o = 0, le = e;
    for (;; ++o) {
      const int d = S.voff[ord_of[o] + 1] - S.voff[ord_of[o]], cnt = d + d * (d + 1) / 2;
      if (le < cnt) break;
      le -= cnt;
    }
    const int d = S.voff[ord_of[o] + 1] - S.voff[ord_of[o]];
    int a2, b2;
    if (le < d) { a2 = le; b2 = 255; }
    else {
      le -= d;
      a2 = 0;
      while (le > a2) { le -= a2 + 1; ++a2; }
      b2 = le;
    }
    htab[e] = (o << 16) | (a2 << 8) | b2;
  }
  __syncthreads();
  for (int iter = 0; iter < 100 && n_ord > 0; ++iter) {
    zero_acc();
    __syncthreads();
    for (int t0 = 0; t0 < n_src; t0 += T) {
      load_tile(t0);
      const bool live = gs[pt] == gs[pt];
      for (int o = sub; o < n_ord; o += R) {
        const int j = ord_of[o], C = S.ncat[j];
        double r1 = 0.0, r0 = 0.0, D1 = 0.0, D0 = 0.0;
        if (live) {
          const int c = (int)ys[pt * p + j];
          const VState s = var_state(S, th, j, xs + pt * Kp, ys[pt * p + j]);
          double pi = ndiff(s.a, s.b);
          if (!(pi > 1e-300)) pi = 1e-300;
          const double p1 = c < C - 1 ? nphi(s.a) : 0.0, p0 = c > 0 ? nphi(s.b) : 0.0;
          r1 = p1 / pi; r0 = p0 / pi; D1 = -s.a * r1; D0 = -s.b * r0;
        }
        double* q = st + (pt * p + o) * 4;
        q[0] = r1; q[1] = r0; q[2] = D1; q[3] = D0;
      }
      __syncthreads();
      chunked(n_h,
              [&](int e, int t) -> double {
                const int code = htab[e], o = code >> 16, a = (code >> 8) & 255, b = code & 255;
                const int j = ord_of[o], C1 = S.ncat[j] - 1;
                const double* q = st + (t * p + o) * 4;
                const int c = (int)ys[t * p + j];
                const double xa = a >= C1 ? xs[t * Kp + S.xidx[j][a - C1]] : 0.0;
                const double j1a = a < C1 ? (a == c ? 1.0 : 0.0) : -xa, j0a = a < C1 ? (a == c - 1 ? 1.0 : 0.0) : -xa;
                const double ga = q[0] * j1a - q[1] * j0a;
                if (b == 255) return ga;
                const double xb = b >= C1 ? xs[t * Kp + S.xidx[j][b - C1]] : 0.0;
                const double j1b = b < C1 ? (b == c ? 1.0 : 0.0) : -xb, j0b = b < C1 ? (b == c - 1 ? 1.0 : 0.0) : -xb;
                const double gb = q[0] * j1b - q[1] * j0b;
                return ga * gb - (q[2] * j1a * j1b - q[3] * j0a * j0b);
              },
              [&](int e, double v) {
                const int code = htab[e], o = code >> 16, a = (code >> 8) & 255, b = code & 255;
                const int d = S.voff[ord_of[o] + 1] - S.voff[ord_of[o]];
                if (b == 255) acc[hoff[o] + a] += v;
                else {
                  acc[hoff[o] + d + a * d + b] += v;
                  if (a != b) acc[hoff[o] + d + b * d + a] += v;
                }
              });
      __syncthreads();
    }
    if (tid == 0) flag_sh = 0;
    __syncthreads();
    if (tid < n_ord) {
      const int j = ord_of[tid], d = S.voff[j + 1] - S.voff[j];
      double* gv = acc + hoff[tid];
      double* H = gv + d;
      if (!thread_chol_solve(H, gv, d)) atomicExch(&flag_sh, 2);
      else {
        double mx = 0.0;
        for (int a = 0; a < d; ++a) { th[S.voff[j] + a] += gv[a]; mx = fmax(mx, fabs(gv[a])); }
        if (mx >= 1e-6) atomicMax(&flag_sh, 1);   // quadratic convergence: the error left is the square of this step
      }
    }
    __syncthreads();
    const int f = flag_sh;
    __syncthreads();
    diag[4] += 1.0;
    if (f == 2) { finish(GWSEM_CATCH_ACOV_NOT_PD, -1); return; }
    if (f == 0) break;
  }
  lap(1);

  // ---------------- phase C: one correlation per pair, Newton on the score
  // Work split between the R threads of a person: polychoric pairs (bivariate normal integrals) and the cheap pairs
  // are dealt out separately.  Per pair and iteration the Gauss-Legendre nodes of the bivariate normal integral
  // (sin t, 1 / (2 cos^2 t), weight * asin(rho) / 2; they depend on rho only) are tabulated once for the CTA.
  const unsigned char* powner = S.powner;
  double* gtab = sm + L.gtab;
  auto fill_gl_tables = [&]() {
    for (int e = tid; e < np * kGlMax; e += NT) {
      const int k = e / kGlMax, i = e % kGlMax;
      const Pair pr = pair_of(k, p);
      if (S.ncat[pr.a] == 0 || S.ncat[pr.b] == 0) continue;
      const double rho = th[ex.n1 + k], ar = fabs(rho), asr = asin(rho);
      // Genz (2004): 6 / 12 / 20 nodes reach 1e-15 in this form for |rho| < 0.3 / 0.75 / 0.925; 32 beyond
      const int rule = ar < 0.3 ? 0 : (ar < 0.75 ? 1 : (ar < 0.925 ? 2 : 3));
      int off = 0;
      for (int r2 = 0; r2 < rule; ++r2) off += 2 * ex.gl_n[r2];
      const int gln = ex.gl_n[rule];
      double* tb = gtab + k * kGlStride;
      if (i == 0) tb[0] = (double)gln;
      if (i < gln) {
        double st2, ct2;
        sincos(0.5 * asr * (ex.gl[off + i] + 1.0), &st2, &ct2);
        tb[1 + 3 * i] = st2;
        tb[2 + 3 * i] = 1.0 / (2.0 * ct2 * ct2);
        tb[3 + 3 * i] = 0.5 * asr * ex.gl[off + gln + i];
      }
    }
  };
  for (int k = tid; k < np; k += NT) rho_ok[k] = 0.0;
  __syncthreads();
  for (int iter = 0; iter < 60 && np > 0; ++iter) {
    zero_acc();
    fill_gl_tables();
    __syncthreads();
    for (int t0 = 0; t0 < n_src; t0 += T) {
      load_tile(t0);
      const bool live = gs[pt] == gs[pt];
      for (int k = 0; k < np; ++k) {
        if (powner[k] != sub) continue;
        double psi = 0.0, dpsi = 0.0;
        if (live && rho_ok[k] == 0.0) {
          const Pair pr = pair_of(k, p);
          const VState va = var_state(S, th, pr.a, xs + pt * Kp, ys[pt * p + pr.a]);
          const VState vb = var_state(S, th, pr.b, xs + pt * Kp, ys[pt * p + pr.b]);
          const PairEval ev = pair_eval(S, pr.a, pr.b, va, vb, (int)ys[pt * p + pr.a], (int)ys[pt * p + pr.b],
                                        th[ex.n1 + k], gtab + k * kGlStride, false);
          psi = ev.psi; dpsi = ev.dpsi;
        }
        ps[(pt * np + k) * 4] = psi;
        ps[(pt * np + k) * 4 + 1] = dpsi;
      }
      __syncthreads();
      for (int e = tid; e < 2 * np; e += NT) {
        const int k = e >> 1, w = e & 1;
        double v = 0.0;
        for (int t = 0; t < T; ++t) v += ps[(t * np + k) * 4 + w];
        acc[e] += v;
      }
      __syncthreads();
    }
    if (tid == 0) flag_sh = 0;
    __syncthreads();
    for (int k = tid; k < np; k += NT) {
      if (rho_ok[k] != 0.0) continue;
      const double G = acc[2 * k], dG = acc[2 * k + 1], r = th[ex.n1 + k];
      double step = dG < 0.0 ? -G / dG : (G > 0.0 ? 0.1 : -0.1);
      step = fmax(-0.3, fmin(0.3, step));
      const double next = fmax(-0.99, fmin(0.99, r + step));
      th[ex.n1 + k] = next;
      if (fabs(next - r) < 1e-6) rho_ok[k] = 1.0;   // Newton: the error left is the square of this step
      else atomicMax(&flag_sh, 1);
    }
    __syncthreads();
    const int f = flag_sh;
    __syncthreads();
    diag[5] += 1.0;
    if (f == 0) break;
  }
  lap(2);

  // ---------------- phase D: score rows at the estimates; B = sum sc sc' (lower triangle), A21 rows
  const int nB = nt * (nt + 1) / 2;
  fill_gl_tables();
  int* etab = reinterpret_cast<int*>(sm + L.etab);
  for (int e = tid; e < nB; e += NT) {
    int b = 0, rem = e;
    while (rem >= nt - b) { rem -= nt - b; ++b; }
    etab[e] = ((b + rem) << 16) | b;   // (a >= b)
  }
  int* atab = etab + nB;
  int n_a21 = 0;
  for (int k = 0; k < np; ++k) {
    const Pair pr = pair_of(k, p);
    n_a21 += (S.voff[pr.a + 1] - S.voff[pr.a]) + (S.voff[pr.b + 1] - S.voff[pr.b]);
  }
  for (int e = tid; e < n_a21; e += NT) {
    int k = 0, le = e;
    for (;; ++k) {
      const Pair pr = pair_of(k, p);
      const int cnt = (S.voff[pr.a + 1] - S.voff[pr.a]) + (S.voff[pr.b + 1] - S.voff[pr.b]);
      if (le < cnt) break;
      le -= cnt;
    }
    const Pair pr = pair_of(k, p);
    const int da = S.voff[pr.a + 1] - S.voff[pr.a];
    atab[e] = (k << 16) | ((le < da ? 1 : 0) << 15) | (le < da ? le : le - da);
  }
  zero_acc();
  __syncthreads();
  for (int t0 = 0; t0 < n_src; t0 += T) {
    load_tile(t0);
    const bool live = gs[pt] == gs[pt];
    double* srow = sc + pt * nt;
    for (int e = sub; e < nt; e += R) srow[e] = 0.0;
    for (int e = sub; e < 4 * np; e += R) ps[pt * np * 4 + e] = 0.0;
    __syncthreads();
    if (live) {
      const double* xr = xs + pt * Kp;
      for (int j = sub; j < p; j += R) {
        const double yv = ys[pt * p + j];
        const VState vj = var_state(S, th, j, xr, yv);
        const int o = S.voff[j], nx = S.nx[j];
        if (S.ncat[j] == 0) {
          const double var = th[o + 1 + nx], e = vj.b;
          srow[o] = e / var;
          for (int m = 0; m < nx; ++m) srow[o + 1 + m] = xr[S.xidx[j][m]] * e / var;
          srow[o + 1 + nx] = (e * e - var) / (2.0 * var * var);
        } else {
          const int C = S.ncat[j], c = (int)yv;
          double pi = ndiff(vj.a, vj.b);
          if (!(pi > 1e-300)) pi = 1e-300;
          const double r1 = c < C - 1 ? nphi(vj.a) / pi : 0.0, r0 = c > 0 ? nphi(vj.b) / pi : 0.0;
          if (c < C - 1) srow[o + c] = r1;
          if (c > 0) srow[o + c - 1] -= r0;
          for (int m = 0; m < nx; ++m) srow[o + C - 1 + m] = xr[S.xidx[j][m]] * (r0 - r1);
        }
      }
      for (int k = 0; k < np; ++k) {
        if (powner[k] != sub) continue;
        const Pair pr = pair_of(k, p);
        const VState va = var_state(S, th, pr.a, xr, ys[pt * p + pr.a]), vb = var_state(S, th, pr.b, xr, ys[pt * p + pr.b]);
        const PairEval ev = pair_eval(S, pr.a, pr.b, va, vb, (int)ys[pt * p + pr.a], (int)ys[pt * p + pr.b],
                                      th[ex.n1 + k], gtab + k * kGlStride, true);
        srow[ex.n1 + k] = ev.psi;
        double* q = ps + (pt * np + k) * 4;
        auto side = [&](int v, double d0, double d1, double* out) {
          if (S.ncat[v] > 0) { out[0] = ev.psi * d0; out[1] = ev.psi * d1; }
          else {
            const double var = th[S.voff[v] + 1 + S.nx[v]], u = (v == pr.a ? va : vb).a;
            out[0] = ev.psi * (-d0 * rsqrt(var));
            out[1] = ev.psi * (-1.0 / (2.0 * var) - d0 * u / (2.0 * var));
          }
        };
        side(pr.a, ev.a0, ev.a1, q);
        side(pr.b, ev.b0, ev.b1, q + 2);
      }
    }
    __syncthreads();
    for (int e = tid; e < nB; e += NT) {
      const int a = etab[e] >> 16, b = etab[e] & 0xffff;
      double v = 0.0;
      for (int t = 0; t < T; ++t) v = fma(sc[t * nt + a], sc[t * nt + b], v);
      acc[e] += v;
    }
    // A21 entries: pair k, then the block of a, then the block of b (atab: pair << 16 | first side << 15 | local index)
    __syncthreads();
    chunked(n_a21,
            [&](int e, int t) -> double {
              const int code = atab[e], k = code >> 16, first = (code >> 15) & 1, ll = code & 0x7fff;
              const Pair pr = pair_of(k, p);
              const int v = first ? pr.a : pr.b;
              const double* q = ps + (t * np + k) * 4 + (first ? 0 : 2);
              return side_entry(S, v, ll, q[0], q[1], ys[t * p + v], xs + t * Kp);
            },
            [&](int e, double v) { acc[nB + e] += v; });
    __syncthreads();
  }
  lap(3);
  if (tid == 0)
    for (int e = 0; e < 8; ++e) rec[ex.fo_B + nt * nt + e] = diag[e];
  // ---------------- record
  for (int e = tid; e < nt; e += NT) rec[ex.fo_theta + e] = th[e];
  for (int e = tid; e < nB; e += NT) {
    const int a = etab[e] >> 16, b = etab[e] & 0xffff;
    rec[ex.fo_B + a * nt + b] = acc[e];
    rec[ex.fo_B + b * nt + a] = acc[e];
  }
  {
    int base = nB;
    for (int k = 0; k < np; ++k) {
      const Pair pr = pair_of(k, p);
      const int da = S.voff[pr.a + 1] - S.voff[pr.a], db = S.voff[pr.b + 1] - S.voff[pr.b];
      for (int l = tid; l < da + db; l += NT) {
        const int c = l < da ? S.voff[pr.a] + l : S.voff[pr.b] + (l - da);
        rec[ex.fo_a21 + k * ex.n1 + c] = acc[base + l];
      }
      base += da + db;
    }
  }
}

}  // namespace

size_t marg_exact_smem_bytes(const ExactProgram& ex) { return (size_t)ex_layout(ex).total * sizeof(double); }

void launch_marg_exact(gwsem_engine* e, const ExactProgram& ex, const uint8_t* d_rows, int64_t pitch, int n_snps, int n_src,
                       const uint8_t* d_inc8, const int32_t* d_counts, int plink1, const double* d_geno, int n_pad,
                       double* d_records) {
  if (n_snps <= 0) return;
  const size_t smem = marg_exact_smem_bytes(ex);
  if (smem > 227 * 1024) fail("marginals model too large for the per-SNP statistics kernel (%zu bytes of scratch)", smem);
  GW_CUDA(cudaFuncSetAttribute(marg_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  e->prof_begin(kFamMargExact);
  // two threads per person of a tile share the per-person work (the bivariate normal integrals dominate it)
  const int threads = ex.T <= 256 ? 2 * ex.T : ex.T;
  marg_exact_kernel<<<n_snps, threads, smem, e->stream>>>(ex, d_rows, pitch, n_src, d_inc8, d_counts, plink1, d_geno, n_pad,
                                                       d_records);
  GW_CUDA(cudaGetLastError());
  e->prof_end(kFamMargExact);
  e->launches++;
}

}  // namespace gwsem
