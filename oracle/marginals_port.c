/* TEST INFRASTRUCTURE ONLY (tests/ and bench.py's CPU-baseline legs; the product never links it).
 *
 * Plain-C restatement of oracle/marginals_oracle.py::fit_snp_marginals -- one row of the reference's GWAS loop for a
 * WLS "marginals" model (mxFitFunctionWLS(allContinuousMethod = "marginals"), reference R/model.R:5, 346-365,
 * 403-418, 435-436; the arithmetic is OpenMx's omxData::estimateObservedStats, restated from Muthen 1984 / Olsson 1979
 * in the outer-product form of lavaan's muthen1984()): stage-1 OLS / ordered probit per manifest variable, stage-2
 * Pearson / polyserial / polychoric per pair, Gamma = A^-1 B A^-T with the delta method, RAM-implied statistics,
 * damped Gauss-Newton with a central-difference Jacobian, (J' W J)^-1.  Same formulas, same iteration rules, same
 * order of operations where it matters; it exists so that the CPU baseline of bench.py is compiled code (-O3) rather
 * than numpy, and is checked against the numpy oracle in tests/test_oracle_marginals_port.py.
 * Scope: manifest variables = columns given by the caller (the snp column included), exogenous covariates = columns
 * given by the caller; rows with a NaN anywhere are dropped (naAction = 'omit').  */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZMAX 12.0
#define MAXV 12     /* manifest variables */
#define MAXK 12     /* covariates */
#define MAXB 32     /* stage-1 parameters of one variable */
#define MAXT 160    /* theta entries */
#define MAXS 160    /* statistics */
#define MAXP 96     /* model parameters */
#define MAXNV 32    /* RAM variables */

typedef struct {
  int n, p, K;                 /* people, manifest variables, exogenous covariates */
  const double* col[MAXV];     /* manifest columns (values or 0-based codes), NaN = missing */
  int kind[MAXV];              /* 0 continuous, 1 ordinal */
  int n_cat[MAXV];
  const double* x[MAXK];       /* covariate columns */
  int exo_free[MAXV][MAXK];
  /* RAM model */
  int nv, nm, n_cells, P;
  const double* cells;         /* n_cells x 5: matrix (0 A, 1 S, 2 M), row, col, parameter or -1, value */
  const double* par_start;
  const double* par_lbound;    /* NaN = none */
  int n_thr;
  const double* thr;           /* n_thr x 4: manifest, k, parameter or -1, value */
  int exo_latent[MAXK];        /* RAM index of each covariate's latent */
  double min_variance;
  int snp_is_exo;              /* the snp column is an exogenous covariate (index of it in x, else -1) */
} PortModel;

typedef struct {
  double est[MAXP];
  double vcov[MAXP * MAXP];
  double fit;
  int n, iterations, status;   /* status: 0 OK, 1 iteration limit, 2 nonzero gradient, 3 not convex, -1 NA */
  int catch_code;              /* 0 none, 1 min variance (catch_var = manifest index, -2 = snp as covariate), 2 acov not positive definite */
  int catch_var;
} PortResult;

/* ------------------------------------------------------------------------------------------ small dense algebra */
static double ndtr(double z) { return 0.5 * erfc(-z * 0.7071067811865476); }
static double phi(double z) { return exp(-0.5 * z * z) * 0.3989422804014327; }
static double ndtri(double p) {   /* inverse of ndtr by Newton from a logistic start (used for threshold starts only) */
  double z = 0.0;
  if (p <= 0.0) return -ZMAX;
  if (p >= 1.0) return ZMAX;
  z = log(p / (1.0 - p)) * 0.6;
  for (int it = 0; it < 100; ++it) {
    const double f = ndtr(z) - p, d = phi(z);
    const double step = f / (d > 1e-300 ? d : 1e-300);
    z -= step;
    if (fabs(step) < 1e-14) break;
  }
  return z;
}
/* solve a x = b (n x n, row-major; a and b destroyed); returns 0 when singular */
static int solve(double* a, double* b, int n, int nrhs) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabs(a[r * n + c]) > fabs(a[piv * n + c])) piv = r;
    if (!(fabs(a[piv * n + c]) > 0.0)) return 0;
    if (piv != c) {
      for (int k = 0; k < n; ++k) { double t = a[c * n + k]; a[c * n + k] = a[piv * n + k]; a[piv * n + k] = t; }
      for (int k = 0; k < nrhs; ++k) { double t = b[c * nrhs + k]; b[c * nrhs + k] = b[piv * nrhs + k]; b[piv * nrhs + k] = t; }
    }
    const double inv = 1.0 / a[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = a[r * n + c] * inv;
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) a[r * n + k] -= f * a[c * n + k];
      for (int k = 0; k < nrhs; ++k) b[r * nrhs + k] -= f * b[c * nrhs + k];
    }
  }
  for (int c = n - 1; c >= 0; --c)
    for (int k = 0; k < nrhs; ++k) {
      double v = b[c * nrhs + k];
      for (int j = c + 1; j < n; ++j) v -= a[c * n + j] * b[j * nrhs + k];
      b[c * nrhs + k] = v / a[c * n + c];
    }
  return 1;
}
static int invert(const double* a, double* out, int n) {
  double* w = (double*)malloc(sizeof(double) * n * n);
  memcpy(w, a, sizeof(double) * n * n);
  for (int i = 0; i < n * n; ++i) out[i] = 0.0;
  for (int i = 0; i < n; ++i) out[i * n + i] = 1.0;
  const int ok = solve(w, out, n, n);
  free(w);
  return ok;
}
/* Cholesky with the oracle's pivot rule (a pivot below 1e-10 of its scale counts as zero) */
static int cholesky_checked(const double* a, int n, const double* scale) {
  double* l = (double*)malloc(sizeof(double) * n * n);
  memcpy(l, a, sizeof(double) * n * n);
  int ok = 1;
  for (int j = 0; j < n && ok; ++j) {
    double d = l[j * n + j];
    const double d0 = scale ? scale[j] : d;
    for (int k = 0; k < j; ++k) d -= l[j * n + k] * l[j * n + k];
    if (!(d > 1e-10 * fabs(d0)) || !isfinite(d)) { ok = 0; break; }
    const double dj = sqrt(d);
    l[j * n + j] = dj;
    for (int i = j + 1; i < n; ++i) {
      double v = l[i * n + j];
      for (int k = 0; k < j; ++k) v -= l[i * n + k] * l[j * n + k];
      l[i * n + j] = v / dj;
    }
  }
  free(l);
  return ok;
}

/* 32-point Gauss-Legendre on [-1, 1] */
static double GLX[32], GLW[32];
static int gl_ready = 0;
static void gl_init(void) {
  if (gl_ready) return;
  const int n = 32;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-16) break;
    }
    GLX[i] = -z; GLX[n - 1 - i] = z;
    GLW[i] = GLW[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  gl_ready = 1;
}
/* P(X <= h, Y <= k), Drezner & Wesolowsky / Genz form; st / ct2 = sin, cos^2 at the nodes on [0, asin rho] */
typedef struct { double a, st[32], ct2[32]; } BvnRule;
static void bvn_rule(double rho, BvnRule* r) {
  gl_init();
  r->a = asin(rho);
  for (int i = 0; i < 32; ++i) {
    const double t = 0.5 * r->a * (GLX[i] + 1.0);
    r->st[i] = sin(t);
    const double c = cos(t);
    r->ct2[i] = c * c;
  }
}
static double bvn_cdf(double h, double k, const BvnRule* r) {
  double acc = 0.0;
  const double hk2 = h * h + k * k, hk = 2.0 * h * k;
  for (int i = 0; i < 32; ++i) acc += GLW[i] * exp(-(hk2 - hk * r->st[i]) / (2.0 * r->ct2[i]));
  return ndtr(h) * ndtr(k) + (0.5 * r->a * acc) / (2.0 * M_PI);
}
static double bvn_pdf(double h, double k, double rho) {
  const double s2 = 1.0 - rho * rho;
  return exp(-(h * h - 2.0 * rho * h * k + k * k) / (2.0 * s2)) / (2.0 * M_PI * sqrt(s2));
}

/* ------------------------------------------------------------------------------------------------ stage 1 */
typedef struct {
  int kind, C, nx, npar, n;
  const double* X[MAXK];     /* free covariates of this variable */
  /* continuous */
  double beta[MAXK + 1], var, sd;
  double* u;                 /* standardised residual */
  double* e;
  /* ordinal */
  int* codes;
  double th[MAXB], gamma[MAXK];
  double *z1, *z0, *p1, *p0, *pi, *resid;
  double* sc;                /* n x npar scores */
} Var;

static void ord_eval(Var* v, const double* par) {
  const int C = v->C, K = v->nx, n = v->n;
  double th[MAXB + 2];
  th[0] = -ZMAX;
  for (int c = 0; c < C - 1; ++c) th[c + 1] = par[c];
  th[C] = ZMAX;
  for (int i = 0; i < n; ++i) {
    double eta = 0.0;
    for (int k = 0; k < K; ++k) eta += v->X[k][i] * par[C - 1 + k];
    const int c = v->codes[i];
    v->z1[i] = th[c + 1] - eta;
    v->z0[i] = th[c] - eta;
    v->p1[i] = c < C - 1 ? phi(v->z1[i]) : 0.0;
    v->p0[i] = c > 0 ? phi(v->z0[i]) : 0.0;
    v->pi[i] = ndtr(v->z1[i]) - ndtr(v->z0[i]);
    v->resid[i] = (v->p0[i] - v->p1[i]) / v->pi[i];
    double* s = v->sc + (size_t)i * v->npar;
    for (int k = 0; k < v->npar; ++k) s[k] = 0.0;
    if (c < C - 1) s[c] = v->p1[i] / v->pi[i];
    if (c > 0) s[c - 1] -= v->p0[i] / v->pi[i];
    for (int k = 0; k < K; ++k) s[C - 1 + k] = v->X[k][i] * v->resid[i];
  }
}
static int fit_ord(Var* v) {
  const int C = v->C, K = v->nx, n = v->n, np = v->npar;
  double par[MAXB], cnt[MAXB + 1];
  for (int c = 0; c < C; ++c) cnt[c] = 0.0;
  for (int i = 0; i < n; ++i) cnt[v->codes[i]] += 1.0;
  double cum = 0.0;
  for (int c = 0; c < C - 1; ++c) {
    cum += cnt[c];
    double q = cum / n;
    q = q < 1e-10 ? 1e-10 : (q > 1 - 1e-10 ? 1 - 1e-10 : q);
    par[c] = ndtri(q);
  }
  for (int k = 0; k < K; ++k) par[C - 1 + k] = 0.0;
  double g[MAXB], H[MAXB * MAXB], J1[MAXB], J0[MAXB], gr[MAXB];
  for (int it = 0; it < 100; ++it) {
    ord_eval(v, par);
    for (int a = 0; a < np; ++a) g[a] = 0.0;
    for (int a = 0; a < np * np; ++a) H[a] = 0.0;
    for (int i = 0; i < n; ++i) {
      const double* s = v->sc + (size_t)i * np;
      for (int a = 0; a < np; ++a) g[a] += s[a];
      /* minus the exact second derivative of the log-likelihood */
      const int c = v->codes[i];
      for (int a = 0; a < np; ++a) J1[a] = J0[a] = 0.0;
      if (c < C - 1) J1[c] = 1.0;
      if (c > 0) J0[c - 1] = 1.0;
      for (int k = 0; k < K; ++k) { J1[C - 1 + k] = -v->X[k][i]; J0[C - 1 + k] = -v->X[k][i]; }
      const double ip = 1.0 / v->pi[i];
      const double d1 = -v->z1[i] * v->p1[i] * ip, d0 = -v->z0[i] * v->p0[i] * ip;
      for (int a = 0; a < np; ++a) gr[a] = (v->p1[i] * J1[a] - v->p0[i] * J0[a]) * ip;
      for (int a = 0; a < np; ++a)
        for (int b = 0; b < np; ++b) H[a * np + b] += gr[a] * gr[b] - (J1[a] * d1 * J1[b] - J0[a] * d0 * J0[b]);
    }
    if (!solve(H, g, np, 1)) return 0;
    double mx = 0.0;
    for (int a = 0; a < np; ++a) { par[a] += g[a]; if (fabs(g[a]) > mx) mx = fabs(g[a]); }
    if (!isfinite(mx)) return 0;
    if (mx < 1e-12) break;
  }
  ord_eval(v, par);
  for (int c = 0; c < C - 1; ++c) v->th[c] = par[c];
  for (int k = 0; k < K; ++k) v->gamma[k] = par[C - 1 + k];
  return 1;
}
static int fit_cont(Var* v, const double* y) {
  const int K = v->nx, n = v->n, q = K + 1;
  double XtX[(MAXK + 1) * (MAXK + 1)], Xty[MAXK + 1], row[MAXK + 1];
  for (int a = 0; a < q * q; ++a) XtX[a] = 0.0;
  for (int a = 0; a < q; ++a) Xty[a] = 0.0;
  for (int i = 0; i < n; ++i) {
    row[0] = 1.0;
    for (int k = 0; k < K; ++k) row[1 + k] = v->X[k][i];
    for (int a = 0; a < q; ++a) { Xty[a] += row[a] * y[i]; for (int b = 0; b < q; ++b) XtX[a * q + b] += row[a] * row[b]; }
  }
  if (!solve(XtX, Xty, q, 1)) return 0;
  for (int a = 0; a < q; ++a) v->beta[a] = Xty[a];
  double ss = 0.0;
  for (int i = 0; i < n; ++i) {
    double f = v->beta[0];
    for (int k = 0; k < K; ++k) f += v->X[k][i] * v->beta[1 + k];
    v->e[i] = y[i] - f;
    ss += v->e[i] * v->e[i];
  }
  v->var = ss / n;
  v->sd = sqrt(v->var);
  for (int i = 0; i < n; ++i) {
    v->u[i] = v->e[i] / v->sd;
    double* s = v->sc + (size_t)i * v->npar;
    s[0] = v->e[i] / v->var;
    for (int k = 0; k < K; ++k) s[1 + k] = v->X[k][i] * v->e[i] / v->var;
    s[1 + K] = (v->e[i] * v->e[i] - v->var) / (2.0 * v->var * v->var);
  }
  return isfinite(v->var) && v->var > 0.0;
}

/* ------------------------------------------------------------------------------------------------ stage 2 */
typedef struct { const Var *a, *b; int type; } Pair;   /* type 0 pearson (a, b continuous), 1 polyserial (a continuous, b ordinal), 2 polychoric */

/* sum over people of the score of rho, and optionally the per-person score (sc) */
static double pair_score(const Pair* pr, double r, double* sc) {
  const Var *a = pr->a, *b = pr->b;
  const int n = a->n;
  double tot = 0.0;
  if (pr->type == 0) {
    const double s2 = 1.0 - r * r;
    for (int i = 0; i < n; ++i) {
      const double u = a->u[i], v = b->u[i];
      const double s = (r * s2 + u * v * (1.0 + r * r) - r * (u * u + v * v)) / (s2 * s2);
      if (sc) sc[i] = s;
      tot += s;
    }
  } else if (pr->type == 1) {
    const double s = sqrt(1.0 - r * r), s3 = s * s * s;
    for (int i = 0; i < n; ++i) {
      const double u = a->u[i];
      const double a1 = (b->z1[i] - r * u) / s, a0 = (b->z0[i] - r * u) / s;
      double pi = ndtr(a1) - ndtr(a0);
      if (pi < 1e-300) pi = 1e-300;
      const double f1 = b->codes[i] < b->C - 1 ? phi(a1) : 0.0, f0 = b->codes[i] > 0 ? phi(a0) : 0.0;
      const double v = (f1 * (r * b->z1[i] - u) / s3 - f0 * (r * b->z0[i] - u) / s3) / pi;
      if (sc) sc[i] = v;
      tot += v;
    }
  } else {
    BvnRule rule;
    bvn_rule(r, &rule);
    for (int i = 0; i < n; ++i) {
      const double az1 = a->z1[i], az0 = a->z0[i], bz1 = b->z1[i], bz0 = b->z0[i];
      double P = bvn_cdf(az1, bz1, &rule) - bvn_cdf(az0, bz1, &rule) - bvn_cdf(az1, bz0, &rule) + bvn_cdf(az0, bz0, &rule);
      if (P < 1e-300) P = 1e-300;
      const int au = a->codes[i] < a->C - 1, al = a->codes[i] > 0, bu = b->codes[i] < b->C - 1, bl = b->codes[i] > 0;
      const double d = (au && bu ? bvn_pdf(az1, bz1, r) : 0.0) - (al && bu ? bvn_pdf(az0, bz1, r) : 0.0) -
                       (au && bl ? bvn_pdf(az1, bz0, r) : 0.0) + (al && bl ? bvn_pdf(az0, bz0, r) : 0.0);
      const double v = d / P;
      if (sc) sc[i] = v;
      tot += v;
    }
  }
  return tot;
}
static double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static double newton_rho(const Pair* pr, double rho0) {
  double rho = rho0, g = pair_score(pr, rho, NULL);
  for (int it = 0; it < 100; ++it) {
    const double h = 1e-6, hi = rho + h < 0.999 ? rho + h : 0.999, lo = rho - h > -0.999 ? rho - h : -0.999;
    const double gp = (pair_score(pr, hi, NULL) - pair_score(pr, lo, NULL)) / (hi - lo);
    const double step = gp < 0.0 ? -g / gp : (g > 0 ? 0.1 : (g < 0 ? -0.1 : 0.0));
    double nw = clip(rho + step, -0.99, 0.99);
    double gn = pair_score(pr, nw, NULL);
    int k = 0;
    while (fabs(gn) > fabs(g) && k < 30) { nw = rho + (nw - rho) / 2.0; gn = pair_score(pr, nw, NULL); ++k; }
    if (fabs(nw - rho) < 1e-13) { rho = nw; break; }
    rho = nw;
    g = gn;
  }
  return rho;
}
/* d loglik_i / d (stage-1 parameters) of a continuous variable from d/du */
static void cont_chain(const Var* c, int i, double dl_du, double* out) {
  const double u = c->u[i];
  out[0] = -dl_du / c->sd;
  for (int k = 0; k < c->nx; ++k) out[1 + k] = c->X[k][i] * (-dl_du / c->sd);
  out[1 + c->nx] = -1.0 / (2.0 * c->var) + dl_du * (-u / (2.0 * c->var));
}
static void ord_chain(const Var* o, int i, double dz1, double dz0, double* out) {
  for (int k = 0; k < o->npar; ++k) out[k] = 0.0;
  const int c = o->codes[i];
  if (c < o->C - 1) out[c] = dz1;
  if (c > 0) out[c - 1] += dz0;
  for (int k = 0; k < o->nx; ++k) out[o->C - 1 + k] = -o->X[k][i] * (dz1 + dz0);
}
/* rho, per-person score sc[n], and the A21 rows sum_i sc_i * d loglik_i / d (stage-1 parameters of a), (of b) */
static double fit_pair(const Pair* pr, double* sc, double* a21a, double* a21b) {
  const Var *a = pr->a, *b = pr->b;
  const int n = a->n;
  double rho0 = 0.0;
  if (pr->type == 0) { for (int i = 0; i < n; ++i) rho0 += a->u[i] * b->u[i]; rho0 /= n; }
  const double rho = newton_rho(pr, rho0);
  pair_score(pr, rho, sc);
  for (int k = 0; k < a->npar; ++k) a21a[k] = 0.0;
  for (int k = 0; k < b->npar; ++k) a21b[k] = 0.0;
  double da[MAXB], db[MAXB];
  const double s2 = 1.0 - rho * rho, s = sqrt(s2);
  BvnRule rule;
  if (pr->type == 2) bvn_rule(rho, &rule);
  for (int i = 0; i < n; ++i) {
    if (pr->type == 0) {
      const double u = a->u[i], v = b->u[i];
      cont_chain(a, i, -(u - rho * v) / s2, da);
      cont_chain(b, i, -(v - rho * u) / s2, db);
    } else if (pr->type == 1) {
      const double u = a->u[i];
      const double a1 = (b->z1[i] - rho * u) / s, a0 = (b->z0[i] - rho * u) / s;
      double pi = ndtr(a1) - ndtr(a0);
      if (pi < 1e-300) pi = 1e-300;
      const double f1 = b->codes[i] < b->C - 1 ? phi(a1) : 0.0, f0 = b->codes[i] > 0 ? phi(a0) : 0.0;
      cont_chain(a, i, -u - (rho / s) * (f1 - f0) / pi, da);
      ord_chain(b, i, f1 / (s * pi), -f0 / (s * pi), db);
    } else {
      const double az1 = a->z1[i], az0 = a->z0[i], bz1 = b->z1[i], bz0 = b->z0[i];
      double P = bvn_cdf(az1, bz1, &rule) - bvn_cdf(az0, bz1, &rule) - bvn_cdf(az1, bz0, &rule) + bvn_cdf(az0, bz0, &rule);
      if (P < 1e-300) P = 1e-300;
      const int au = a->codes[i] < a->C - 1, al = a->codes[i] > 0, bu = b->codes[i] < b->C - 1, bl = b->codes[i] > 0;
#define DZ(za, zb1, zb0, mask) ((mask) ? phi(za) * (ndtr(((zb1) - rho * (za)) / s) - ndtr(((zb0) - rho * (za)) / s)) : 0.0)
      ord_chain(a, i, DZ(az1, bz1, bz0, au) / P, -DZ(az0, bz1, bz0, al) / P, da);
      ord_chain(b, i, DZ(bz1, az1, az0, bu) / P, -DZ(bz0, az1, az0, bl) / P, db);
#undef DZ
    }
    for (int k = 0; k < a->npar; ++k) a21a[k] += sc[i] * da[k];
    for (int k = 0; k < b->npar; ++k) a21b[k] += sc[i] * db[k];
  }
  return rho;
}

/* ------------------------------------------------------------------------------------------------- the model */
typedef struct { int kind, j, sub; } Stat;   /* 0 mean, 1 threshold, 2 slope (sub = covariate), 3 variance, 4 covariance (j > sub) */

static void implied(const PortModel* m, const double* theta, const Stat* layout, int ns, double* out) {
  const int nv = m->nv, nm = m->nm;
  double A[MAXNV * MAXNV], S[MAXNV * MAXNV], M[MAXNV], Z[MAXNV * MAXNV], T[MAXNV * MAXNV], Cm[MAXNV * MAXNV], mu[MAXNV];
  for (int e = 0; e < nv * nv; ++e) A[e] = S[e] = 0.0;
  for (int e = 0; e < nv; ++e) M[e] = 0.0;
  for (int c = 0; c < m->n_cells; ++c) {
    const double* x = m->cells + 5 * (size_t)c;
    const int mat = (int)x[0], r = (int)x[1], cc = (int)x[2], par = (int)x[3];
    const double val = par >= 0 ? theta[par] : x[4];
    if (mat == 0) A[r * nv + cc] = val;
    else if (mat == 1) S[r * nv + cc] = val;
    else M[cc] = val;
  }
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) T[i * nv + j] = (i == j ? 1.0 : 0.0) - A[i * nv + j];
  invert(T, Z, nv);
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double v = 0.0;
      for (int k = 0; k < nv; ++k) v += Z[i * nv + k] * S[k * nv + j];
      T[i * nv + j] = v;
    }
  for (int i = 0; i < nm; ++i)
    for (int j = 0; j < nm; ++j) {
      double v = 0.0;
      for (int k = 0; k < nv; ++k) v += T[i * nv + k] * Z[j * nv + k];
      Cm[i * nv + j] = v;
    }
  for (int i = 0; i < nm; ++i) {
    double v = 0.0;
    for (int k = 0; k < nv; ++k) v += Z[i * nv + k] * M[k];
    mu[i] = v;
  }
  for (int e = 0; e < ns; ++e) {
    const int j = layout[e].j, sub = layout[e].sub;
    const double sdj = m->kind[j] ? sqrt(Cm[j * nv + j]) : 1.0;
    double v;
    switch (layout[e].kind) {
      case 0: v = mu[j]; break;
      case 1: {
        double tau = 0.0;
        for (int t = 0; t < m->n_thr; ++t)
          if ((int)m->thr[4 * t] == j && (int)m->thr[4 * t + 1] == sub) tau = (int)m->thr[4 * t + 2] >= 0 ? theta[(int)m->thr[4 * t + 2]] : m->thr[4 * t + 3];
        v = (tau - mu[j]) / sdj;
        break;
      }
      case 2: v = Z[j * nv + m->exo_latent[sub]] / sdj; break;
      case 3: v = Cm[j * nv + j]; break;
      default: {
        const double sds = m->kind[sub] ? sqrt(Cm[sub * nv + sub]) : 1.0;
        v = Cm[j * nv + sub] / (sdj * sds);
      }
    }
    out[e] = v;
  }
}

static void jacobian(const PortModel* m, const double* theta, const Stat* layout, int ns, double* J) {
  const int P = m->P;
  double tp[MAXP], tm[MAXP], sp[MAXS], sm[MAXS];
  for (int k = 0; k < P; ++k) {
    const double h = 1e-6 * (fabs(theta[k]) > 1.0 ? fabs(theta[k]) : 1.0);
    memcpy(tp, theta, sizeof(double) * P);
    memcpy(tm, theta, sizeof(double) * P);
    tp[k] += h;
    tm[k] -= h;
    implied(m, tp, layout, ns, sp);
    implied(m, tm, layout, ns, sm);
    for (int e = 0; e < ns; ++e) J[e * P + k] = (sp[e] - sm[e]) / (2.0 * h);
  }
}

static void* zalloc(size_t bytes) { void* p = calloc(1, bytes ? bytes : 1); return p; }

int gwsem_port_fit_marginals(const PortModel* m, PortResult* out) {
  const int p = m->p, K = m->K, P = m->P;
  for (int k = 0; k < P; ++k) out->est[k] = m->par_start[k];
  for (int k = 0; k < P * P; ++k) out->vcov[k] = NAN;
  out->fit = NAN; out->n = 0; out->iterations = 0; out->status = -1; out->catch_code = 0; out->catch_var = -1;
  if (p > MAXV || K > MAXK || P > MAXP || m->nv > MAXNV) return 1;
  /* naAction = 'omit' */
  int* keep = (int*)malloc(sizeof(int) * (m->n > 0 ? m->n : 1));
  int n = 0;
  for (int i = 0; i < m->n; ++i) {
    int ok = 1;
    for (int j = 0; j < p && ok; ++j) ok = isfinite(m->col[j][i]);
    for (int k = 0; k < K && ok; ++k) ok = isfinite(m->x[k][i]);
    if (ok) keep[n++] = i;
  }
  out->n = n;
  double* cols = (double*)zalloc(sizeof(double) * (size_t)n * (p + K));
  for (int j = 0; j < p; ++j) for (int i = 0; i < n; ++i) cols[(size_t)j * n + i] = m->col[j][keep[i]];
  for (int k = 0; k < K; ++k) for (int i = 0; i < n; ++i) cols[(size_t)(p + k) * n + i] = m->x[k][keep[i]];
  free(keep);
  int rc = 0;
  Var vars[MAXV];
  memset(vars, 0, sizeof(vars));
  double *sc2 = NULL, *A = NULL, *B = NULL, *acov = NULL, *Hm = NULL, *nacov = NULL, *W = NULL, *SC = NULL;
#define SAMPLE_VAR(ptr, res) { double mean = 0.0, ss = 0.0; for (int i = 0; i < n; ++i) mean += (ptr)[i]; mean /= n; \
    for (int i = 0; i < n; ++i) ss += ((ptr)[i] - mean) * ((ptr)[i] - mean); res = n > 1 ? ss / (n - 1) : 0.0; }
  if (m->snp_is_exo >= 0) {
    double v; SAMPLE_VAR(cols + (size_t)(p + m->snp_is_exo) * n, v);
    if (!(v >= m->min_variance)) { out->catch_code = 1; out->catch_var = -2; goto done; }
  }
  for (int j = 0; j < p; ++j)
    if (!m->kind[j]) {
      double v; SAMPLE_VAR(cols + (size_t)j * n, v);
      if (!(v >= m->min_variance)) { out->catch_code = 1; out->catch_var = j; goto done; }
    }
  /* ---- stage 1 */
  int off[MAXV + 1];
  off[0] = 0;
  for (int j = 0; j < p; ++j) {
    Var* v = &vars[j];
    v->kind = m->kind[j]; v->C = m->n_cat[j]; v->n = n; v->nx = 0;
    for (int k = 0; k < K; ++k) if (m->exo_free[j][k]) v->X[v->nx++] = cols + (size_t)(p + k) * n;
    v->npar = v->kind ? v->C - 1 + v->nx : v->nx + 2;
    v->sc = (double*)zalloc(sizeof(double) * (size_t)n * v->npar);
    int ok;
    if (v->kind) {
      v->codes = (int*)zalloc(sizeof(int) * n);
      for (int i = 0; i < n; ++i) v->codes[i] = (int)cols[(size_t)j * n + i];
      v->z1 = (double*)zalloc(sizeof(double) * n); v->z0 = (double*)zalloc(sizeof(double) * n);
      v->p1 = (double*)zalloc(sizeof(double) * n); v->p0 = (double*)zalloc(sizeof(double) * n);
      v->pi = (double*)zalloc(sizeof(double) * n); v->resid = (double*)zalloc(sizeof(double) * n);
      ok = fit_ord(v);
    } else {
      v->u = (double*)zalloc(sizeof(double) * n); v->e = (double*)zalloc(sizeof(double) * n);
      ok = fit_cont(v, cols + (size_t)j * n);
    }
    if (!ok) { out->catch_code = 2; goto done; }
    off[j + 1] = off[j] + v->npar;
  }
  {
    const int n1 = off[p], npairs = p * (p - 1) / 2, nt = n1 + npairs;
    if (nt > MAXT) { rc = 1; goto done; }
    /* ---- stage 2 */
    sc2 = (double*)zalloc(sizeof(double) * (size_t)n * npairs);
    A = (double*)zalloc(sizeof(double) * nt * nt);
    double rho[MAXV * MAXV];
    int pi_[MAXV * MAXV], pj_[MAXV * MAXV];
    int k = 0;
    for (int j = 0; j < p; ++j)
      for (int i = j + 1; i < p; ++i, ++k) {
        pi_[k] = i; pj_[k] = j;
        Pair pr;
        double a21x[MAXB], a21y[MAXB];
        const Var *a = &vars[i], *b = &vars[j];
        if (!a->kind && !b->kind) { pr.a = a; pr.b = b; pr.type = 0; }
        else if (!a->kind) { pr.a = a; pr.b = b; pr.type = 1; }
        else if (!b->kind) { pr.a = b; pr.b = a; pr.type = 1; }
        else { pr.a = a; pr.b = b; pr.type = 2; }
        rho[k] = fit_pair(&pr, sc2 + (size_t)k * n, a21x, a21y);
        const int ia = (int)(pr.a - vars), ib = (int)(pr.b - vars);
        for (int q = 0; q < pr.a->npar; ++q) A[(n1 + k) * nt + off[ia] + q] = a21x[q];
        for (int q = 0; q < pr.b->npar; ++q) A[(n1 + k) * nt + off[ib] + q] = a21y[q];
      }
    /* ---- Gamma of theta: A^-1 B A^-T */
    SC = (double*)zalloc(sizeof(double) * (size_t)n * nt);
    for (int j = 0; j < p; ++j)
      for (int i = 0; i < n; ++i)
        for (int q = 0; q < vars[j].npar; ++q) SC[(size_t)i * nt + off[j] + q] = vars[j].sc[(size_t)i * vars[j].npar + q];
    for (int q = 0; q < npairs; ++q)
      for (int i = 0; i < n; ++i) SC[(size_t)i * nt + n1 + q] = sc2[(size_t)q * n + i];
    B = (double*)zalloc(sizeof(double) * nt * nt);
    for (int i = 0; i < n; ++i) {
      const double* s = SC + (size_t)i * nt;
      for (int a = 0; a < nt; ++a) { const double sa = s[a]; if (sa == 0.0) continue; for (int b = 0; b < nt; ++b) B[a * nt + b] += sa * s[b]; }
    }
    for (int j = 0; j < p; ++j)
      for (int a = off[j]; a < off[j + 1]; ++a)
        for (int b = off[j]; b < off[j + 1]; ++b) A[a * nt + b] = B[a * nt + b];
    for (int q = 0; q < npairs; ++q) A[(n1 + q) * nt + n1 + q] = B[(n1 + q) * nt + n1 + q];
    double* Ainv = (double*)zalloc(sizeof(double) * nt * nt);
    if (!invert(A, Ainv, nt)) { free(Ainv); out->catch_code = 2; goto done; }
    double* T = (double*)zalloc(sizeof(double) * nt * nt);
    acov = (double*)zalloc(sizeof(double) * nt * nt);
    for (int a = 0; a < nt; ++a) for (int b = 0; b < nt; ++b) { double v = 0.0; for (int q = 0; q < nt; ++q) v += Ainv[a * nt + q] * B[q * nt + b]; T[a * nt + b] = v; }
    for (int a = 0; a < nt; ++a) for (int b = 0; b < nt; ++b) { double v = 0.0; for (int q = 0; q < nt; ++q) v += T[a * nt + q] * Ainv[b * nt + q]; acov[a * nt + b] = v; }
    free(T); free(Ainv);
    /* ---- statistics and the delta method */
    Stat layout[MAXS];
    double s[MAXS];
    int ns = 0;
    Hm = (double*)zalloc(sizeof(double) * MAXS * nt);
#define PUSH(kd, jj, sb, val) { if (ns >= MAXS) { rc = 1; goto done; } layout[ns].kind = kd; layout[ns].j = jj; layout[ns].sub = sb; s[ns] = val; ++ns; }
    for (int j = 0; j < p; ++j) {
      if (!vars[j].kind) { PUSH(0, j, 0, vars[j].beta[0]); Hm[(ns - 1) * nt + off[j]] = 1.0; }
      else for (int c = 0; c < vars[j].C - 1; ++c) { PUSH(1, j, c, vars[j].th[c]); Hm[(ns - 1) * nt + off[j] + c] = 1.0; }
    }
    for (int j = 0; j < p; ++j) {
      const int base = off[j] + (vars[j].kind ? vars[j].C - 1 : 1);
      int mth = 0;
      for (int kk = 0; kk < K; ++kk)
        if (m->exo_free[j][kk]) {
          PUSH(2, j, kk, vars[j].kind ? vars[j].gamma[mth] : vars[j].beta[1 + mth]);
          Hm[(ns - 1) * nt + base + mth] = 1.0;
          ++mth;
        }
    }
    for (int j = 0; j < p; ++j)
      if (!vars[j].kind) { PUSH(3, j, j, vars[j].var); Hm[(ns - 1) * nt + off[j + 1] - 1] = 1.0; }
    for (int q = 0; q < npairs; ++q) {
      const int i = pi_[q], j = pj_[q];
      const double sa = vars[i].kind ? 1.0 : vars[i].sd, sb = vars[j].kind ? 1.0 : vars[j].sd;
      PUSH(4, i, j, rho[q] * sa * sb);
      double* row = Hm + (size_t)(ns - 1) * nt;
      row[n1 + q] = sa * sb;
      if (!vars[i].kind) row[off[i + 1] - 1] = rho[q] * sb / (2.0 * sa);
      if (!vars[j].kind) row[off[j + 1] - 1] = rho[q] * sa / (2.0 * sb);
    }
    nacov = (double*)zalloc(sizeof(double) * ns * ns);
    {
      double* HT = (double*)zalloc(sizeof(double) * ns * nt);
      for (int a = 0; a < ns; ++a) for (int b = 0; b < nt; ++b) { double v = 0.0; for (int q = 0; q < nt; ++q) v += Hm[a * nt + q] * acov[q * nt + b]; HT[a * nt + b] = v; }
      for (int a = 0; a < ns; ++a) for (int b = 0; b < ns; ++b) { double v = 0.0; for (int q = 0; q < nt; ++q) v += HT[a * nt + q] * Hm[b * nt + q]; nacov[a * ns + b] = n * v; }
      free(HT);
    }
    W = (double*)zalloc(sizeof(double) * ns * ns);
    if (!invert(nacov, W, ns) || !cholesky_checked(nacov, ns, NULL)) { out->catch_code = 2; goto done; }
    for (int e = 0; e < ns * ns; ++e) W[e] *= n;
    /* ---- damped Gauss-Newton (fit_oracle.wls_fit) */
    double theta[MAXP], cand[MAXP], r[MAXS], rcv[MAXS], sig[MAXS], g[MAXP], delta[MAXP];
    double* J = (double*)zalloc(sizeof(double) * ns * P);
    double* H = (double*)zalloc(sizeof(double) * P * P);
    double* Hc = (double*)zalloc(sizeof(double) * P * P);
    double* WJ = (double*)zalloc(sizeof(double) * ns * P);
    memcpy(theta, m->par_start, sizeof(double) * P);
#define OBJ(th, res, F) { implied(m, th, layout, ns, sig); for (int e = 0; e < ns; ++e) res[e] = s[e] - sig[e]; F = 0.0; \
    for (int a = 0; a < ns; ++a) { double v = 0.0; for (int b = 0; b < ns; ++b) v += W[a * ns + b] * res[b]; F += res[a] * v; } }
    double F, Fc = 0.0, lam = 1e-4;
    OBJ(theta, r, F);
    int status = 1, it = 0;
    for (it = 1; it <= 200; ++it) {
      jacobian(m, theta, layout, ns, J);
      for (int a = 0; a < ns; ++a) for (int k2 = 0; k2 < P; ++k2) { double v = 0.0; for (int b = 0; b < ns; ++b) v += W[a * ns + b] * J[b * P + k2]; WJ[a * P + k2] = v; }
      for (int k2 = 0; k2 < P; ++k2) { double v = 0.0; for (int a = 0; a < ns; ++a) v += WJ[a * P + k2] * r[a]; g[k2] = v; }
      for (int a = 0; a < P; ++a) for (int b = 0; b < P; ++b) { double v = 0.0; for (int e = 0; e < ns; ++e) v += J[e * P + a] * WJ[e * P + b]; H[a * P + b] = v; }
      int improved = 0;
      for (int attempt = 0; attempt < 30; ++attempt) {
        memcpy(Hc, H, sizeof(double) * P * P);
        for (int a = 0; a < P; ++a) Hc[a * P + a] += lam * (H[a * P + a] > 1e-12 ? H[a * P + a] : 1e-12);
        memcpy(delta, g, sizeof(double) * P);
        if (!solve(Hc, delta, P, 1)) { lam *= 10; continue; }
        for (int k2 = 0; k2 < P; ++k2) {
          double v = theta[k2] + delta[k2];
          const double lb = m->par_lbound[k2];
          if (lb == lb && v < lb) v = lb;
          cand[k2] = v;
        }
        OBJ(cand, rcv, Fc);
        if (isfinite(Fc) && Fc <= F) { improved = 1; break; }
        lam *= 10;
      }
      if (!improved) {
        double gmax = 0.0;
        for (int k2 = 0; k2 < P; ++k2) if (fabs(g[k2]) > gmax) gmax = fabs(g[k2]);
        status = gmax < 1e-6 * (1 + F) ? 0 : 2;
        break;
      }
      double step = 0.0;
      for (int k2 = 0; k2 < P; ++k2) { const double q = fabs(cand[k2] - theta[k2]) / (fabs(theta[k2]) + 1e-3); if (q > step) step = q; }
      const double dF = F - Fc;
      memcpy(theta, cand, sizeof(double) * P);
      memcpy(r, rcv, sizeof(double) * ns);
      F = Fc;
      lam = lam / 10 > 1e-12 ? lam / 10 : 1e-12;
      if (step < 1e-10 || dF <= 1e-14 * (1 + F)) { status = 0; break; }
    }
    if (it > 200) it = 200;
    jacobian(m, theta, layout, ns, J);
    for (int a = 0; a < ns; ++a) for (int k2 = 0; k2 < P; ++k2) { double v = 0.0; for (int b = 0; b < ns; ++b) v += W[a * ns + b] * J[b * P + k2]; WJ[a * P + k2] = v; }
    for (int a = 0; a < P; ++a) for (int b = 0; b < P; ++b) { double v = 0.0; for (int e = 0; e < ns; ++e) v += J[e * P + a] * WJ[e * P + b]; H[a * P + b] = v; }
    for (int a = 0; a < P; ++a) for (int b = 0; b < a; ++b) { const double v = 0.5 * (H[a * P + b] + H[b * P + a]); H[a * P + b] = H[b * P + a] = v; }
    if (cholesky_checked(H, P, NULL) && invert(H, out->vcov, P)) { /* vcov filled */ }
    else { for (int k2 = 0; k2 < P * P; ++k2) out->vcov[k2] = NAN; if (status == 0) status = 3; }
    memcpy(out->est, theta, sizeof(double) * P);
    out->fit = F; out->iterations = it; out->status = status;
    free(J); free(H); free(Hc); free(WJ);
  }
done:
  for (int j = 0; j < p; ++j) {
    free(vars[j].sc); free(vars[j].codes); free(vars[j].z1); free(vars[j].z0); free(vars[j].p1); free(vars[j].p0);
    free(vars[j].pi); free(vars[j].resid); free(vars[j].u); free(vars[j].e);
  }
  free(sc2); free(A); free(B); free(acov); free(Hm); free(nacov); free(W); free(SC); free(cols);
  return rc;
}

size_t gwsem_port_sizeof_model(void) { return sizeof(PortModel); }
size_t gwsem_port_sizeof_result(void) { return sizeof(PortResult); }
