"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the ML / FIML row of the per-SNP loop
(reference call sites R/model.R:6 (mxFitFunctionML), 166-173
(GradientDescent, NumericDeriv, StandardError, HessianQuality), 403-418 (exogenous covariates as
definition variables), 421-437 (raw data, gxe product columns)).  Only tests/, smoke() and the
cpu_baseline leg of bench.py may import this.

PARITY WEAKLY PINNED: the arithmetic lives in OpenMx (not in the reference tree, no R here).  The one reference
assertion on an ML fit, tests/testthat/test-model.R:68-79 (the P values of buildItem under ML correlate 1 +- 1e-3
with those under WLS, which are pinned absolutely), is reproduced in tests/test_reference_pins.py on the R-RNG replay
of that test's phenotypes; no reference test pins a multi-item ML estimate.  What is restated is full-information
maximum likelihood for continuous manifest variables as OpenMx
defines it:
  -2 lnL = sum_i [ p_i log 2 pi + log |Sigma_i| + (y_i - mu_i)' Sigma_i^-1 (y_i - mu_i) ]
with Sigma_i, mu_i the model-implied moments filtered to the variables observed for row i, and
mu_i = mu + slopes x_i for the definition-variable covariates x_i.  People whose genotype call is
missing keep contributing through their items (snp and its gxe products are the missing
variables); rows with a missing item or covariate are left out on both sides (the device path does
the same; the reference would keep the observed items of such rows).
Parameter covariance: 2 H^-1, H the Hessian of -2 lnL at the estimate (mxComputeNumericDeriv +
mxComputeStandardError), here by central differences of the analytic gradient."""
import numpy as np

from fit_oracle import FlatRAM

LOG2PI = np.log(2 * np.pi)


def _implied(ram, flat, theta):
    A, S, M = ram.matrices(theta)
    nv, nm = ram.nv, ram.nm
    Z = np.linalg.inv(np.eye(nv) - A)
    C = (Z @ S @ Z.T)[:nm, :nm]
    mu = (Z @ M)[:nm]
    exo = [nm + flat["latent"].index(c) for c in flat["exo_pred"]]
    B = Z[:nm][:, exo] if exo else np.zeros((nm, 0))
    return C, np.column_stack([mu, B])          # Sigma, coefficient matrix on w = [1, x]


def _implied_jac(ram, flat, theta):
    """d Sigma / d theta_k [P, nm, nm] and d Coef / d theta_k [P, nm, 1 + K]."""
    A, S, M = ram.matrices(theta)
    nv, nm = ram.nv, ram.nm
    Z = np.linalg.inv(np.eye(nv) - A)
    C = Z @ S @ Z.T
    zm = Z @ M
    exo = [nm + flat["latent"].index(c) for c in flat["exo_pred"]]
    dS = np.zeros((ram.P, nm, nm))
    dC = np.zeros((ram.P, nm, 1 + len(exo)))
    for m, r, c, p, v in ram.cells:
        if p < 0:
            continue
        if m == 0:
            t = np.outer(Z[:, r], C[c, :])
            dS[p] += (t + t.T)[:nm, :nm]
            dC[p][:, 0] += Z[:nm, r] * zm[c]
            for k, x in enumerate(exo):
                dC[p][:, 1 + k] += Z[:nm, r] * Z[c, x]
        elif m == 1:
            dS[p] += np.outer(Z[:, r], Z[:, c])[:nm, :nm]
        else:
            dC[p][:, 0] += Z[:nm, c]
    return dS, dC


class Group:
    """People sharing a pattern of observed manifest variables, as augmented moments of [w, y]."""
    def __init__(self, idx, Y, W):
        self.idx = np.asarray(idx)
        self.n = Y.shape[0]
        self.Mww = W.T @ W
        self.Mwy = W.T @ Y
        self.Myy = Y.T @ Y


def neg2ll(ram, flat, theta, groups, grad=False):
    Sig, Cf = _implied(ram, flat, theta)
    if grad:
        dS, dC = _implied_jac(ram, flat, theta)
        g = np.zeros(ram.P)
    f = 0.0
    for gr in groups:
        if gr.n == 0:
            continue
        ix = gr.idx
        S = Sig[np.ix_(ix, ix)]
        Cg = Cf[ix]
        try:
            L = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            return (np.inf, np.zeros(ram.P)) if grad else np.inf
        G = np.linalg.inv(S)
        R = gr.Mwy - gr.Mww @ Cg.T                  # [1+K, p]
        T = gr.Myy - Cg @ gr.Mwy - gr.Mwy.T @ Cg.T + Cg @ gr.Mww @ Cg.T
        f += gr.n * (len(ix) * LOG2PI + 2 * np.log(np.diag(L)).sum()) + np.trace(G @ T)
        if grad:
            Q = gr.n * G - G @ T @ G
            U = G @ R.T                             # [p, 1+K]
            for k in range(ram.P):
                g[k] += np.sum(Q * dS[k][np.ix_(ix, ix)]) - 2 * np.sum(U * dC[k][ix])
    return (f, g) if grad else f


def fit_snp_ml(flat, pheno_cols, snp):
    """One row of the GWAS loop for a fitfun='ML' model with continuous manifest variables."""
    from scipy.optimize import minimize
    ram = FlatRAM(flat)
    P, names = ram.P, flat["manifest"]
    out = dict(est=np.array(flat["par_start"], float), vcov=np.full((P, P), np.nan), status="NA", catch1="",
               fit=np.nan, n=0, iterations=0)
    snp = np.asarray(snp, float)
    cols = []
    for m in names:
        if m == "snp":
            cols.append(snp)
        elif m.startswith("snp_"):
            with np.errstate(invalid="ignore"):
                cols.append(snp * pheno_cols[m[4:]])
        else:
            cols.append(np.asarray(pheno_cols[m], float))
    Y = np.column_stack(cols)
    # buildItem's default for ML is exogenous=TRUE: snp itself is then a definition variable
    # (R/model.R:560-576) and a person without a call has no likelihood contribution at all
    snp_exo = "snp" in flat["exo_pred"]
    X = (np.column_stack([snp if c == "snp" else pheno_cols[c] for c in flat["exo_pred"]])
         if flat["exo_pred"] else np.zeros((len(snp), 0)))
    dep = np.array([m == "snp" or m.startswith("snp_") for m in names])
    xs = np.array([c != "snp" for c in flat["exo_pred"]], bool)
    static_ok = np.isfinite(Y[:, ~dep]).all(1) & np.isfinite(X[:, xs]).all(1)
    for m in names:   # a moderator column must be complete too
        if m.startswith("snp_"):
            static_ok &= np.isfinite(pheno_cols[m[4:]])
    obs = static_ok & np.isfinite(snp)
    mis = static_ok & ~np.isfinite(snp) & (not snp_exo)
    out["n"] = int(obs.sum())
    for j, m in enumerate(names):
        if not (np.var(Y[obs, j], ddof=1) >= flat["min_variance"]):
            out["catch1"] = f"{flat['name']}.data: '{m}' has observed variance less than {flat['min_variance']:g}"
            return out
    W = np.column_stack([np.ones(len(snp)), np.where(np.isfinite(X), X, 0.0)])
    if snp_exo and not (np.var(snp[obs], ddof=1) >= flat["min_variance"]):
        out["catch1"] = f"{flat['name']}.data: 'snp' has observed variance less than {flat['min_variance']:g}"
        return out
    groups = [Group(np.arange(len(names)), Y[obs], W[obs]),
              Group(np.nonzero(~dep)[0], Y[mis][:, ~dep], W[mis])]
    lb = np.where(np.isnan(flat["par_lbound"]), -np.inf, flat["par_lbound"])
    bounds = [(l if np.isfinite(l) else None, None) for l in lb]
    fun = lambda t: neg2ll(ram, flat, t, groups, grad=True)
    n_all = groups[0].n + groups[1].n
    def scaled(t):
        f, g = fun(t)
        return (f / n_all, g / n_all) if np.isfinite(f) else (1e300, np.zeros(P))
    res = minimize(scaled, np.array(flat["par_start"], float), jac=True,
                   method="L-BFGS-B", bounds=bounds, options=dict(maxiter=2000, ftol=1e-15, gtol=1e-12, maxcor=40))
    theta = res.x
    # polish with Newton steps on the finite-difference Hessian of the analytic gradient
    def hessian(t):
        H = np.zeros((P, P))
        for k in range(P):
            h = 1e-5 * max(1.0, abs(t[k]))
            e = np.zeros(P); e[k] = h
            H[:, k] = (fun(t + e)[1] - fun(t - e)[1]) / (2 * h)
        return 0.5 * (H + H.T)
    for _ in range(5):
        f, g = fun(theta)
        H = hessian(theta)
        free = theta > lb + 1e-12
        step = np.zeros(P)
        try:
            step[free] = np.linalg.solve(H[np.ix_(free, free)], g[free])
        except np.linalg.LinAlgError:
            break
        cand = np.maximum(theta - step, lb)
        if neg2ll(ram, flat, cand, groups) <= f:
            theta = cand
        if np.max(np.abs(step) / (np.abs(theta) + 1e-3)) < 1e-9:
            break
    f, g = fun(theta)
    H = hessian(theta)
    gmax = np.max(np.abs(g[theta > lb + 1e-12])) if np.any(theta > lb + 1e-12) else 0.0
    status = "OK" if gmax < 1e-4 * n_all else "nonzero gradient/red"
    try:    # a Hessian that is not positive definite at the solution: OpenMx statusCode 5, no covariance
        from fit_oracle import cholesky_checked
        cholesky_checked(H)
        vcov = 2 * np.linalg.inv(H)
    except np.linalg.LinAlgError:
        vcov = np.full((P, P), np.nan)
        if status == "OK":
            status = "not convex/red"
    out.update(est=theta, vcov=vcov, status=status, fit=f, iterations=int(res.nit))
    return out


def neg2ll_rows(flat, theta, pheno_cols, snp):
    """The same -2 lnL, row by row (definition of FIML); for checking the moment form on small data."""
    ram = FlatRAM(flat)
    names = flat["manifest"]
    Sig, Cf = _implied(ram, flat, theta)
    total = 0.0
    for i in range(len(snp)):
        y = []
        for m in names:
            if m == "snp":
                y.append(snp[i])
            elif m.startswith("snp_"):
                y.append(snp[i] * pheno_cols[m[4:]][i])
            else:
                y.append(pheno_cols[m][i])
        y = np.array(y, float)
        w = np.array([1.0] + [pheno_cols[c][i] for c in flat["exo_pred"]])
        dep = np.array([m == "snp" or m.startswith("snp_") for m in names])
        if not (np.isfinite(y[~dep]).all() and np.isfinite(w).all()):
            continue
        ix = np.nonzero(np.isfinite(y))[0]
        r = y[ix] - Cf[ix] @ w
        S = Sig[np.ix_(ix, ix)]
        total += len(ix) * LOG2PI + np.linalg.slogdet(S)[1] + r @ np.linalg.solve(S, r)
    return total
