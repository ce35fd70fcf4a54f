"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64) restatement of the WLS *marginals* path
(ordinal items and/or exogenous covariates): mxFitFunctionWLS(allContinuousMethod="marginals"),
reference R/model.R:5, selected whenever postprocessModel does not switch to cumulants
(R/model.R:488-510); thresholds from setupThresholds (R/model.R:346-365), exogenous covariates
as definition-variable latents with `exoFree` slopes (R/model.R:403-418).

PARITY PINNED BY THE REFERENCE'S KNOWN ANSWERS (round 2): the arithmetic lives in OpenMx (DESCRIPTION:49-51), absent
from the reference tree and from this image, but the reference's tests assert numbers it produced on phenotypes
simulated with R's RNG, which oracle/r_rng.py replays bit for bit.  tests/test_reference_pins.py reproduces
test-model.R:97,103 (P == .0251 / .086, ordinal + continuous depVars with an exogenous snp), :122 (snp_to_F of the
7-item one-factor model, to the three printed decimals), :157 (P == c(.455, .885, .395)), :171 (facCov == 1.17) and
test-formats.R:65-70 (on example.bed: 171 clean / 28 failed rows against 173 / 26 +- 6, fivenum(Z) to 1e-2), each at
the reference's own tolerance except the two noted in DESIGN.md section 3.  Restated from the published three-stage estimator
(Muthen 1984; Olsson 1979; Olsson, Drasgow & Dorans 1982) in the outer-product form used by
lavaan's muthen1984(), which OpenMx's omxData::estimateObservedStats states it follows:
  stage 1  per variable given its exogenous predictors: OLS (continuous; ML variance /N) or
           ordered probit (thresholds + slopes, Newton-Raphson)
  stage 2  per pair, one correlation by ML with stage-1 parameters fixed: Pearson, polyserial,
           polychoric
  Gamma    A^-1 B A^-T from per-person scores, A = [[A11, 0], [A21, A22]] (outer products),
           then the delta method to (means/thresholds, slopes, variances, covariances) where a
           continuous variable contributes sd * rho
  model    RAM moments conditional on the exogenous covariates, ordinal rows standardised by
           the model-implied sd (comment at R/model.R:327-332)
Checked in tests/test_oracle_marginals.py by finite differences of every log-likelihood,
agreement with closed forms, and Monte-Carlo calibration of Gamma.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import numpy as np
from scipy.special import ndtr

from fit_oracle import FlatRAM, cholesky_checked, wls_fit

SQRT2PI = np.sqrt(2 * np.pi)
ZMAX = 12.0  # stands in for +-infinity thresholds


def phi(z):
    return np.exp(-0.5 * z * z) / SQRT2PI


_GL_X, _GL_W = np.polynomial.legendre.leggauss(32)


def bvn_cdf(h, k, rho):
    """P(X <= h, Y <= k), standard bivariate normal (Drezner & Wesolowsky 1990 / Genz 2004 form):
    Phi(h)Phi(k) + 1/(2 pi) * int_0^asin(rho) exp(-(h^2 + k^2 - 2 h k sin t) / (2 cos^2 t)) dt."""
    h = np.asarray(h, float)
    k = np.asarray(k, float)
    a = np.arcsin(rho)
    t = 0.5 * a * (_GL_X + 1.0)                  # nodes on [0, asin rho]
    st, ct2 = np.sin(t), np.cos(t) ** 2
    hh, kk = h[..., None], k[..., None]
    integrand = np.exp(-(hh * hh + kk * kk - 2 * hh * kk * st) / (2 * ct2))
    return ndtr(h) * ndtr(k) + (0.5 * a * (integrand * _GL_W).sum(-1)) / (2 * np.pi)


def bvn_pdf(h, k, rho):
    s2 = 1 - rho * rho
    return np.exp(-(h * h - 2 * rho * h * k + k * k) / (2 * s2)) / (2 * np.pi * np.sqrt(s2))


# ------------------------------------------------------------------------------------ stage 1

class Cont:
    """OLS of y on [1, X]; parameters (intercept, slopes..., variance)."""

    def __init__(self, y, X):
        n = len(y)
        self.X1 = np.column_stack([np.ones(n), X]) if X.shape[1] else np.ones((n, 1))
        self.beta = np.linalg.solve(self.X1.T @ self.X1, self.X1.T @ y)
        self.e = y - self.X1 @ self.beta
        self.var = self.e @ self.e / n
        self.sd = np.sqrt(self.var)
        self.u = self.e / self.sd
        # scores d loglik_i / d (beta, var)
        self.sc_beta = self.X1 * (self.e / self.var)[:, None]
        self.sc_var = (self.e ** 2 - self.var) / (2 * self.var ** 2)
        self.kind = "cont"


class Ord:
    """Ordered probit of codes (0..C-1) on X (no intercept); parameters (thresholds, slopes)."""

    def __init__(self, codes, X, n_cat, max_iter=100):
        self.codes = codes.astype(int)
        self.X = X
        self.C = n_cat
        n, K = X.shape
        cum = np.cumsum(np.bincount(self.codes, minlength=n_cat))[:-1] / n
        from scipy.special import ndtri
        th = ndtri(np.clip(cum, 1e-10, 1 - 1e-10))
        par = np.concatenate([th, np.zeros(K)])
        for _ in range(max_iter):
            self._eval(par)
            g = self.sc.sum(0)
            H = self.sc.T @ self.sc if False else self._hessian(par)
            step = np.linalg.solve(H, g)
            par = par + step
            if np.max(np.abs(step)) < 1e-12:
                break
        self._eval(par)
        self.th, self.gamma = par[:n_cat - 1], par[n_cat - 1:]
        self.kind = "ord"

    def _eval(self, par):
        C, K = self.C, self.X.shape[1]
        th = np.concatenate([[-ZMAX], par[:C - 1], [ZMAX]])
        eta = self.X @ par[C - 1:] if K else np.zeros(len(self.codes))
        self.z1 = th[self.codes + 1] - eta
        self.z0 = th[self.codes] - eta
        self.p1, self.p0 = phi(self.z1) * (self.codes < C - 1), phi(self.z0) * (self.codes > 0)
        self.pi = ndtr(self.z1) - ndtr(self.z0)
        n = len(self.codes)
        sc_th = np.zeros((n, C - 1))
        rows = np.arange(n)
        up = self.codes < C - 1
        sc_th[rows[up], self.codes[up]] = (self.p1 / self.pi)[up]
        lo = self.codes > 0
        sc_th[rows[lo], self.codes[lo] - 1] -= (self.p0 / self.pi)[lo]
        self.resid = (self.p0 - self.p1) / self.pi           # generalised residual E[e | category]
        self.sc = np.column_stack([sc_th, self.X * self.resid[:, None]]) if K else sc_th

    def _hessian(self, par):
        """minus the exact second derivative of the log-likelihood"""
        C, K = self.C, self.X.shape[1]
        n = len(self.codes)
        # d/dz of (p1 - p0)/pi pieces: phi'(z) = -z phi(z)
        d1 = -self.z1 * self.p1                    # d phi(z1)/dz1
        d0 = -self.z0 * self.p0
        # Jacobian of z1, z0 wrt parameters
        J1 = np.zeros((n, C - 1 + K))
        J0 = np.zeros((n, C - 1 + K))
        rows = np.arange(n)
        up = self.codes < C - 1
        J1[rows[up], self.codes[up]] = 1
        lo = self.codes > 0
        J0[rows[lo], self.codes[lo] - 1] = 1
        if K:
            J1[:, C - 1:] = -self.X
            J0[:, C - 1:] = -self.X
        g = (self.p1[:, None] * J1 - self.p0[:, None] * J0)           # d pi / d par
        H = (g / self.pi[:, None]).T @ (g / self.pi[:, None])
        H -= (J1 * (d1 / self.pi)[:, None]).T @ J1 - (J0 * (d0 / self.pi)[:, None]).T @ J0
        return H


# ------------------------------------------------------------------------------------ stage 2

def _newton_rho(loglik_grad, rho0=0.0):
    """1-D Newton / secant on the score with step halving; returns rho."""
    rho = rho0
    g = loglik_grad(rho)
    for _ in range(100):
        h = 1e-6
        gp = (loglik_grad(min(rho + h, 0.999)) - loglik_grad(max(rho - h, -0.999))) / (min(rho + h, 0.999) - max(rho - h, -0.999))
        step = -g / gp if gp < 0 else np.sign(g) * 0.1
        new = np.clip(rho + step, -0.99, 0.99)
        gn = loglik_grad(new)
        k = 0
        while abs(gn) > abs(g) and k < 30:
            new = rho + (new - rho) / 2
            gn = loglik_grad(new)
            k += 1
        if abs(new - rho) < 1e-13:
            rho = new
            break
        rho, g = new, gn
    return float(rho)


def pair_pearson(a, b):
    """Both continuous.  Returns rho, score_rho[n], d loglik_i / d (stage-1 params of a), (of b)."""
    u, v = a.u, b.u

    def score(r):
        s2 = 1 - r * r
        return (r * s2 + u * v * (1 + r * r) - r * (u * u + v * v)) / (s2 * s2)

    rho = _newton_rho(lambda r: score(r).sum(), float(np.mean(u * v)))
    s2 = 1 - rho * rho
    sc = score(rho)
    # d loglik / du and / dv of the bivariate density, then chain to (beta, var) of each variable
    dl_du = -(u - rho * v) / s2
    dl_dv = -(v - rho * u) / s2
    da = np.column_stack([a.X1 * (-dl_du / a.sd)[:, None], -1 / (2 * a.var) + dl_du * (-u / (2 * a.var))])
    db = np.column_stack([b.X1 * (-dl_dv / b.sd)[:, None], -1 / (2 * b.var) + dl_dv * (-v / (2 * b.var))])
    return rho, sc, da, db


def pair_polyserial(c, o):
    """c continuous, o ordinal."""
    u = c.u

    def parts(r):
        s = np.sqrt(1 - r * r)
        a1, a0 = (o.z1 - r * u) / s, (o.z0 - r * u) / s
        pi = ndtr(a1) - ndtr(a0)
        f1, f0 = phi(a1) * (o.codes < o.C - 1), phi(a0) * (o.codes > 0)
        return s, a1, a0, np.maximum(pi, 1e-300), f1, f0

    def score(r):
        s, a1, a0, pi, f1, f0 = parts(r)
        da1 = (r * o.z1 - u) / s ** 3
        da0 = (r * o.z0 - u) / s ** 3
        return (f1 * da1 - f0 * da0) / pi

    rho = _newton_rho(lambda r: score(r).sum(), 0.0)
    s, a1, a0, pi, f1, f0 = parts(rho)
    sc = score(rho)
    dl_dz1 = f1 / (s * pi)
    dl_dz0 = -f0 / (s * pi)
    dl_du = -u - (rho / s) * (f1 - f0) / pi
    dc = np.column_stack([c.X1 * (-dl_du / c.sd)[:, None], -1 / (2 * c.var) + dl_du * (-u / (2 * c.var))])
    do = _ord_chain(o, dl_dz1, dl_dz0)
    return rho, sc, dc, do


def _ord_chain(o, dl_dz1, dl_dz0):
    """d loglik_i / d (thresholds, slopes) of an ordinal variable from d/dz1, d/dz0."""
    n = len(o.codes)
    out = np.zeros((n, o.C - 1 + o.X.shape[1]))
    rows = np.arange(n)
    up = o.codes < o.C - 1
    out[rows[up], o.codes[up]] = dl_dz1[up]
    lo = o.codes > 0
    out[rows[lo], o.codes[lo] - 1] += dl_dz0[lo]
    if o.X.shape[1]:
        out[:, o.C - 1:] = -o.X * (dl_dz1 + dl_dz0)[:, None]
    return out


def pair_polychoric(a, b):
    def prob(r):
        return (bvn_cdf(a.z1, b.z1, r) - bvn_cdf(a.z0, b.z1, r) - bvn_cdf(a.z1, b.z0, r) + bvn_cdf(a.z0, b.z0, r))

    def score(r):
        P = np.maximum(prob(r), 1e-300)
        d = (bvn_pdf(a.z1, b.z1, r) * ((a.codes < a.C - 1) & (b.codes < b.C - 1))
             - bvn_pdf(a.z0, b.z1, r) * ((a.codes > 0) & (b.codes < b.C - 1))
             - bvn_pdf(a.z1, b.z0, r) * ((a.codes < a.C - 1) & (b.codes > 0))
             + bvn_pdf(a.z0, b.z0, r) * ((a.codes > 0) & (b.codes > 0)))
        return d / P

    rho = _newton_rho(lambda r: score(r).sum(), 0.0)
    s = np.sqrt(1 - rho * rho)
    P = np.maximum(prob(rho), 1e-300)
    sc = score(rho)

    def dz(za, zb1, zb0, mask):   # d P / d za for a rectangle edge at za
        return phi(za) * mask * (ndtr((zb1 - rho * za) / s) - ndtr((zb0 - rho * za) / s))

    da = _ord_chain(a, dz(a.z1, b.z1, b.z0, a.codes < a.C - 1) / P, -dz(a.z0, b.z1, b.z0, a.codes > 0) / P)
    db = _ord_chain(b, dz(b.z1, a.z1, a.z0, b.codes < b.C - 1) / P, -dz(b.z0, a.z1, a.z0, b.codes > 0) / P)
    return rho, sc, da, db


# ------------------------------------------------------------------ observed statistics + Gamma

def marginals_stats(columns, kinds, n_cats, X, exo_free):
    """columns: list of arrays (continuous values or 0-based category codes), kinds: 'cont'/'ord',
    X[n, K] exogenous covariates, exo_free[p, K] bool.  Returns dict with the statistic vector in
    the order (per variable: mean | thresholds), (slopes: variable-major over its free covariates),
    (variances of continuous variables), (lower-triangle covariances / correlations, column-major),
    its N-scaled asymptotic covariance NACOV (so that acov(s) = NACOV / n), and a layout description.
    """
    n = len(columns[0])
    p = len(columns)
    K = X.shape[1]
    vars_ = []
    for j in range(p):
        Xj = X[:, exo_free[j]] if K else np.zeros((n, 0))
        vars_.append(Cont(columns[j], Xj) if kinds[j] == "cont" else Ord(columns[j], Xj, n_cats[j]))
    # stage-1 parameter blocks, in theta_1 order: per variable [means/thresholds, slopes, (variance)]
    blocks, sc1 = [], []
    for v in vars_:
        if v.kind == "cont":
            sc = np.column_stack([v.sc_beta, v.sc_var])
        else:
            sc = v.sc
        blocks.append(sc.shape[1])
        sc1.append(sc)
    off = np.concatenate([[0], np.cumsum(blocks)])
    n1 = off[-1]
    pairs = [(i, j) for j in range(p) for i in range(j + 1, p)]
    rho = np.zeros(len(pairs))
    sc2 = np.zeros((n, len(pairs)))
    A21 = np.zeros((len(pairs), n1))
    for k, (i, j) in enumerate(pairs):
        a, b = vars_[i], vars_[j]
        if a.kind == "cont" and b.kind == "cont":
            r, sc, da, db = pair_pearson(a, b)
        elif a.kind == "cont":
            r, sc, da, db = pair_polyserial(a, b)
        elif b.kind == "cont":
            r, sc, db, da = pair_polyserial(b, a)
        else:
            r, sc, da, db = pair_polychoric(a, b)
        rho[k] = r
        sc2[:, k] = sc
        A21[k, off[i]:off[i + 1]] = sc @ da
        A21[k, off[j]:off[j + 1]] = sc @ db
    SC = np.column_stack(sc1 + [sc2])
    B = SC.T @ SC
    A = np.zeros((n1 + len(pairs), n1 + len(pairs)))
    for j in range(p):
        A[off[j]:off[j + 1], off[j]:off[j + 1]] = sc1[j].T @ sc1[j]
    A[n1:, :n1] = A21
    A[n1:, n1:] = np.diag((sc2 ** 2).sum(0))
    Ainv = np.linalg.inv(A)
    acov_theta = Ainv @ B @ Ainv.T           # covariance of (theta_1, rho), not N-scaled
    # ---- delta method to the statistic vector
    theta = np.concatenate([np.concatenate([v.beta, [v.var]]) if v.kind == "cont" else np.concatenate([v.th, v.gamma])
                            for v in vars_] + [rho])
    layout = []      # (kind, var index, sub index)
    rows = []
    nt = len(theta)

    def unit(idx):
        e = np.zeros(nt)
        e[idx] = 1
        return e

    stats = []
    for j, v in enumerate(vars_):                       # means / thresholds
        if v.kind == "cont":
            stats.append(v.beta[0]); rows.append(unit(off[j])); layout.append(("mean", j, 0))
        else:
            for c in range(v.C - 1):
                stats.append(v.th[c]); rows.append(unit(off[j] + c)); layout.append(("thr", j, c))
    for j, v in enumerate(vars_):                       # slopes
        nsl = int(exo_free[j].sum()) if K else 0
        base = off[j] + (1 if v.kind == "cont" else v.C - 1)
        coef = v.beta[1:] if v.kind == "cont" else v.gamma
        ks = np.nonzero(exo_free[j])[0] if K else []
        for m in range(nsl):
            stats.append(coef[m]); rows.append(unit(base + m)); layout.append(("slope", j, int(ks[m])))
    for j, v in enumerate(vars_):                       # variances
        if v.kind == "cont":
            stats.append(v.var); rows.append(unit(off[j + 1] - 1)); layout.append(("var", j, j))
    for k, (i, j) in enumerate(pairs):                  # covariances: rho * sd_i * sd_j (sd = 1 for ordinal)
        a, b = vars_[i], vars_[j]
        sa = a.sd if a.kind == "cont" else 1.0
        sb = b.sd if b.kind == "cont" else 1.0
        stats.append(rho[k] * sa * sb)
        row = unit(n1 + k) * sa * sb
        if a.kind == "cont":
            row[off[i + 1] - 1] = rho[k] * sb / (2 * sa)
        if b.kind == "cont":
            row[off[j + 1] - 1] = rho[k] * sa / (2 * sb)
        rows.append(row)
        layout.append(("cov", i, j))
    Hm = np.array(rows)
    nacov = n * (Hm @ acov_theta @ Hm.T)
    return dict(s=np.array(stats), nacov=nacov, layout=layout, n=n, vars=vars_, rho=rho, pairs=pairs)


# ------------------------------------------------------------------------- model-implied vector

def implied_marginals(ram, flat, theta, layout, ordinal):
    """sigma(theta) in the order of `layout`.  ordinal[j] = True for ordinal manifest j."""
    A, S, M = ram.matrices(theta)
    nv, nm = ram.nv, ram.nm
    Z = np.linalg.inv(np.eye(nv) - A)
    C = (Z @ S @ Z.T)[:nm, :nm]
    mu = (Z @ M)[:nm]
    exo_idx = [flat["manifest"].__len__() + flat["latent"].index(c) for c in flat["exo_pred"]]
    slopes = Z[:nm][:, exo_idx] if exo_idx else np.zeros((nm, 0))
    sd = np.where(ordinal, np.sqrt(np.diag(C)), 1.0)
    thr = {}
    for mi, k, par, val in flat["thresholds"]:
        thr[(int(mi), int(k))] = theta[int(par)] if par >= 0 else val
    out = []
    for kind, j, sub in layout:
        if kind == "mean":
            out.append(mu[j])
        elif kind == "thr":
            out.append((thr[(j, sub)] - mu[j]) / sd[j])
        elif kind == "slope":
            out.append(slopes[j, sub] / sd[j])
        elif kind == "var":
            out.append(C[j, j])
        else:
            out.append(C[j, sub] / (sd[j] * sd[sub]))
    return np.array(out)


def fit_snp_marginals(flat, pheno_cols, snp, n_cats):
    """One row of the GWAS loop for a WLS / marginals model whose manifest variables are `snp`
    (continuous) and items (continuous or ordinal, pheno_cols[item] = values or 0-based codes), with
    exogenous covariates flat['exo_pred'] regressed on as flat['exo_free'] says."""
    ram = FlatRAM(flat)
    P = ram.P
    out = dict(est=np.array(flat["par_start"], float), vcov=np.full((P, P), np.nan), status="NA", catch1="",
               fit=np.nan, n=0, iterations=0)
    names = flat["manifest"]
    def manifest_column(m):   # `snp_<moderator>` as a manifest variable: the gxe algebra column (R/model.R:427-434)
        if m == "snp":
            return snp
        if m.startswith("snp_") and m[4:] in flat["gxe"] and m not in pheno_cols:
            with np.errstate(invalid="ignore"):
                return snp * pheno_cols[m[4:]]
        return pheno_cols[m]
    cols = [manifest_column(m) for m in names]
    # buildItem with an ordered phenotype: `snp` itself is an exogenous predictor (R/model.R:560-576)
    def exo_column(c):   # `snp_<moderator>`: the gxe product as a definition variable (R/model.R:427-434)
        if c == "snp":
            return snp
        if c.startswith("snp_") and c[4:] in flat["gxe"] and c not in pheno_cols:
            with np.errstate(invalid="ignore"):
                return snp * pheno_cols[c[4:]]
        return pheno_cols[c]
    X = (np.column_stack([exo_column(c) for c in flat["exo_pred"]])
         if flat["exo_pred"] else np.zeros((len(snp), 0)))
    keep = np.isfinite(np.column_stack(cols + [X])).all(1)
    cols = [c[keep] for c in cols]
    X = X[keep]
    out["n"] = int(keep.sum())
    kinds = ["ord" if n_cats.get(m, 0) > 1 else "cont" for m in names]
    if "snp" in flat["exo_pred"] and not (np.var(snp[keep], ddof=1) >= flat["min_variance"]):
        out["catch1"] = f"{flat['name']}.data: 'snp' has observed variance less than {flat['min_variance']:g}"
        return out
    for j, m in enumerate(names):
        if kinds[j] == "cont" and not (np.var(cols[j], ddof=1) >= flat["min_variance"]):
            out["catch1"] = f"{flat['name']}.data: '{m}' has observed variance less than {flat['min_variance']:g}"
            return out
    exo_free = flat["exo_free"] if flat["exo_free"] is not None else np.zeros((len(names), 0), bool)
    try:
        st = marginals_stats(cols, kinds, [n_cats.get(m, 0) for m in names], X, exo_free)
    except np.linalg.LinAlgError:  # a stage-1 regression is not identified (e.g. a constant snp as predictor)
        out["catch1"] = f"{flat['name']}.data: asymptotic covariance matrix is not positive definite"
        return out
    n = st["n"]
    try:
        W = n * np.linalg.inv(st["nacov"])
        np.linalg.cholesky(st["nacov"])
    except np.linalg.LinAlgError:
        out["catch1"] = f"{flat['name']}.data: asymptotic covariance matrix is not positive definite"
        return out
    ordinal = np.array([k == "ord" for k in kinds])
    sig = lambda t: implied_marginals(ram, flat, t, st["layout"], ordinal)

    def jac(t):
        J = np.zeros((len(st["s"]), P))
        for k in range(P):
            h = 1e-6 * max(1.0, abs(t[k]))
            e = np.zeros(P)
            e[k] = h
            J[:, k] = (sig(t + e) - sig(t - e)) / (2 * h)
        return J

    theta, F, it, status = wls_fit(st["s"], W, sig, jac, flat["par_start"], flat["par_lbound"])
    J = jac(theta)
    try:    # an information matrix that is not positive definite at the solution: OpenMx statusCode 5, no covariance
        cholesky_checked(J.T @ W @ J)
        vcov = np.linalg.inv(J.T @ W @ J)
    except np.linalg.LinAlgError:
        vcov = np.full((P, P), np.nan)
        if status == "OK":
            status = "not convex/red"
    out.update(est=theta, vcov=vcov, status=status, fit=F, iterations=it, stats=st)
    return out
